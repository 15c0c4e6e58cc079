B="--no-e2e --no-cpu-baseline --no-configs"
for v in x1 x8 "" x16 x20 x24; do
L=python-raytracer_b200/csrc/libsightpy_b200${v:+_$v}.so
SIGHTPY_B200_LIB=$L timeout 300 python bench.py --config stress --spp 4 --steps 2 --warmup 1 $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stress ${v:-x12}', round(d['value']), d['ms_per_step'], [round(x,1) for x in d['config']['level_ms_rank0']], d['frame']['mean_radiance'])"
done
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 300 -k "bvh or stress or pretrace or shadow or mesh" 2>&1 | tail -4
