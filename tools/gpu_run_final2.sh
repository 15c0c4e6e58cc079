# Final single-GPU measurements of round 2 (second session): tests, bench lines of every configuration, reference arm,
# ncu launch lists + per-ray figures, one full ncu capture per kernel family.  Outputs under gpurun_out/r2b_*.
set -x
P=gpurun_out/r2b
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 > ${P}_tests.log 2>&1; echo "tests rc=$?"; tail -3 ${P}_tests.log
SIGHTPY_B200_LIB=python-raytracer_b200/csrc/libsightpy_b200_checked.so timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 > ${P}_tests_checked.log 2>&1; echo "checked tests rc=$?"; tail -3 ${P}_tests_checked.log
S="--width 960 --height 540 --spp 4 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs"
M="gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 200 python bench.py $S > ${P}_small_cornell.json 2>/dev/null
timeout 600 ncu --metrics $M --print-units base --clock-control none --csv --log-file ${P}_launches_cornell.csv python bench.py $S > /dev/null 2>&1; echo "ncu cornell rc=$?"
python tools/ncu_per_ray.py ${P}_launches_cornell.csv ${P}_small_cornell.json cornell profiles/r2b_per_ray.json > /dev/null
for c in example2 example3 example4; do
  S2="--config $c --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --spp 2"
  timeout 200 python bench.py $S2 > ${P}_small_$c.json 2>/dev/null
  timeout 600 ncu --metrics $M --print-units base --clock-control none --csv --log-file ${P}_launches_$c.csv python bench.py $S2 > /dev/null 2>&1; echo "ncu $c rc=$?"
  python tools/ncu_per_ray.py ${P}_launches_$c.csv ${P}_small_$c.json $c profiles/r2b_per_ray.json > /dev/null
done
S5="--config stress --width 1920 --height 1080 --spp 1 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-configs"
timeout 200 python bench.py $S5 > ${P}_small_stress.json 2>/dev/null
timeout 600 ncu --metrics $M --print-units base --clock-control none --csv --log-file ${P}_launches_stress.csv python bench.py $S5 > /dev/null 2>&1; echo "ncu stress rc=$?"
python tools/ncu_per_ray.py ${P}_launches_stress.csv ${P}_small_stress.json stress profiles/r2b_per_ray.json > /dev/null
timeout 500 python bench.py > ${P}_final_bench.json 2> ${P}_final_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference > ${P}_final_reference_arm.json 2> ${P}_final_reference_arm.err; echo "ref rc=$?"
for c in example1 example2 example3 example4; do timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-configs > ${P}_final_bench_$c.json 2> ${P}_final_bench_$c.err; echo "$c rc=$?"; done
timeout 600 python bench.py --config stress --spp 4 --steps 2 --warmup 1 --no-configs > ${P}_final_bench_stress_4spp.json 2> ${P}_final_bench_stress_4spp.err; echo "stress rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sp_warp -s 7 -c 5 -f -o ${P}_final_warp python bench.py $S > /dev/null 2>&1; echo "ncu full warp rc=$?"
python tools/anim_probe.py 300 2>&1 | grep -v proccesing > ${P}_anim.log; cat ${P}_anim.log
for leaf in 2 3 6; do SIGHTPY_BVH_LEAF=$leaf timeout 200 python bench.py --config stress --spp 4 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stress leaf=$leaf', round(d['value']), d['ms_per_step'])"; done
