timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 300 2>&1 | tail -5
B="--no-e2e --no-cpu-baseline --no-configs"
timeout 300 python bench.py --steps 2 --warmup 1 $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cornell', round(d['value']), d['ms_per_step'], d['config']['level_ms_rank0'], d['frame']['mean_radiance'])"
for c in example2 example3 example4; do
timeout 300 python bench.py --config $c --steps 20 --warmup 5 $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$c', round(d['value']), d['ms_per_step'], d['config']['level_ms_rank0'], d['frame']['mean_radiance'])"
done
timeout 300 python bench.py --config stress --spp 4 --steps 2 --warmup 1 $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stress', round(d['value']), d['ms_per_step'], d['config']['level_ms_rank0'], d['frame']['mean_radiance'])"
