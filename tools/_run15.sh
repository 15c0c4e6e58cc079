timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 300 -k "split or example or fuzz or tonemap or frame" 2>&1 | tail -5
B="--no-e2e --no-cpu-baseline --no-configs"
for sp in 0 1 2; do
for c in example1 example2 example3 example4; do
SIGHTPY_SPLIT=$sp timeout 300 python bench.py --config $c --steps 20 --warmup 5 $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('split=$sp $c', round(d['value']), round(d['ms_per_step'],3), [round(x,3) for x in d['config']['level_ms_rank0']], d['frame']['mean_radiance'], d['gpu_launches'])"
done; done
