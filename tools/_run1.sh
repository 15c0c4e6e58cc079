for c in example2 example3 example4; do
for sw in "3 3" "200 100"; do set -- $sw
timeout 300 python bench.py --config $c --steps $1 --warmup $2 --no-e2e --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$c', '$sw', round(d['value']), d['ms_per_step'], d['config']['level_ms_rank0'], d['clocks'])"
done; done
