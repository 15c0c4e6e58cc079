S="--config stress --width 1920 --height 1080 --spp 1 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-configs"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sp_trace -s 7 -c 1 -f -o gpurun_out/r2b_stress_trace python bench.py $S > /dev/null 2>&1; echo "ncu trace rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sp_shadow -s 7 -c 1 -f -o gpurun_out/r2b_stress_shadow python bench.py $S > /dev/null 2>&1; echo "ncu shadow rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sp_level -s 7 -c 1 -f -o gpurun_out/r2b_stress_level python bench.py $S > /dev/null 2>&1; echo "ncu level rc=$?"
python bench.py $S 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['ms_per_step'], d['config']['rays_per_depth_rank0'], d['config']['level_ms_rank0'])"
