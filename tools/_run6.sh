timeout 600 python tools/check_texels.py 2>&1 | grep -v proccesing | tail -30
for c in example2 example3 example4; do
timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$c', round(d['value']), d['ms_per_step'], d['config']['level_ms_rank0'])"
done
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 300 2>&1 | tail -8
