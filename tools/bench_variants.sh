#!/bin/bash
# run bench.py against every library variant in build/variants (development aid)
for lib in build/variants/*.so; do
  SIGHTPY_B200_LIB=$PWD/$lib python bench.py --spp ${SPP:-32} --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms', d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
