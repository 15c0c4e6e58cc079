# Bench lines of the final tree (after the two-chunks-in-flight change): default line + one line per configuration.
P=gpurun_out/r2b
timeout 500 python bench.py > ${P}_final_bench.json 2> ${P}_final_bench.err; echo "bench rc=$?"
for c in example1 example2 example3 example4; do timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-configs > ${P}_final_bench_$c.json 2> ${P}_final_bench_$c.err; echo "$c rc=$?"; done
timeout 600 python bench.py --config stress --spp 4 --steps 2 --warmup 1 --no-configs > ${P}_final_bench_stress_4spp.json 2> ${P}_final_bench_stress_4spp.err; echo "stress rc=$?"
python tools/anim_probe.py 300 2>&1 | grep -v "proccesing\|blurring" > ${P}_anim.log; cat ${P}_anim.log
