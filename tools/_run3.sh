S2="--config example2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --spp 2"
SIGHTPY_PHASE_TIMING=1 SIGHTPY_B200_LIB=python-raytracer_b200/csrc/libsightpy_b200_pt.so python bench.py $S2 2>&1 | grep -v "^{" | tail -40
python bench.py --config example2 --steps 300 --warmup 50 --no-e2e --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['ms_per_step'], d['clocks'])"
