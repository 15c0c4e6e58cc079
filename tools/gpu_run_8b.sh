# 8-GPU measurements of the final tree (gpurun --gpus 8): two-GPU tests, torchrun arm as the driver launches it, the in-process arm, the stress scene
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
P=gpurun_out/r2b
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
timeout 400 $TR bench.py --gpus 8 > ${P}_final_bench_8gpu.json 2> ${P}_final_bench_8gpu.err; echo "torchrun8 rc=$?"
timeout 300 python bench.py --gpus 8 --in-process > ${P}_final_bench_8gpu_inprocess.json 2> ${P}_final_bench_8gpu_inprocess.err; echo "inprocess8 rc=$?"
timeout 300 $TR bench.py --gpus 8 --config stress --steps 2 --warmup 1 --no-configs --no-cpu-baseline > ${P}_final_stress_8gpu_tiles.json 2> ${P}_final_stress_8gpu_tiles.err; echo "stress tiles rc=$?"
timeout 300 $TR bench.py --gpus 8 --config stress --shard samples --steps 2 --warmup 1 --no-configs --no-cpu-baseline > ${P}_final_stress_8gpu_samples.json 2> ${P}_final_stress_8gpu_samples.err; echo "stress samples rc=$?"
for c in example2 example4; do timeout 300 $TR bench.py --gpus 8 --config $c --steps 10 --warmup 3 --no-configs --no-cpu-baseline > ${P}_final_${c}_8gpu.json 2> ${P}_final_${c}_8gpu.err; echo "$c rc=$?"; done
for f in ${P}_final_bench_8gpu.json ${P}_final_bench_8gpu_inprocess.json ${P}_final_stress_8gpu_tiles.json ${P}_final_stress_8gpu_samples.json ${P}_final_example2_8gpu.json ${P}_final_example4_8gpu.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d['value']), d['ms_per_step'], d.get('e2e') and round(d['e2e']['value']), d['frame']['mean_radiance'], d['frame']['sha256'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
