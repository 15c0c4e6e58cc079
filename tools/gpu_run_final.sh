set -x
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 > gpurun_out/r2_tests6.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2_tests6.log
timeout 500 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference > gpurun_out/r2_final_reference_arm.json 2> gpurun_out/r2_final_reference_arm.err; echo "ref rc=$?"
for c in example1 example2 example3 example4; do timeout 300 python bench.py --config $c --no-configs > gpurun_out/r2_final_bench_$c.json 2> gpurun_out/r2_final_bench_$c.err; echo "$c rc=$?"; done
timeout 600 python bench.py --config stress --steps 1 --warmup 1 --no-configs > gpurun_out/r2_final_bench_stress.json 2> gpurun_out/r2_final_bench_stress.err; echo "stress rc=$?"
S="--width 960 --height 540 --spp 4 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs"
M="gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 200 python bench.py $S > gpurun_out/r2_small_cornell.json 2>/dev/null
timeout 600 ncu --metrics $M --print-units base --clock-control none --csv --log-file gpurun_out/r2_launches_cornell.csv python bench.py $S > /dev/null 2>&1; echo "ncu cornell rc=$?"
for c in example2 example4; do
  S2="--config $c --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --spp 2"
  timeout 200 python bench.py $S2 > gpurun_out/r2_small_$c.json 2>/dev/null
  timeout 600 ncu --metrics $M --print-units base --clock-control none --csv --log-file gpurun_out/r2_launches_$c.csv python bench.py $S2 > /dev/null 2>&1; echo "ncu $c rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sp_warp -s 7 -c 5 -f -o gpurun_out/r2_final_warp python bench.py $S > /dev/null 2>&1; echo "ncu full rc=$?"
