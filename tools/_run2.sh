S2="--config example2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --spp 2"
echo "== plain"; python bench.py $S2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['ms_per_step'], d['config']['level_ms_rank0'])"
echo "== phase timing"; SIGHTPY_PHASE_TIMING=1 SIGHTPY_B200_LIB=python-raytracer_b200/csrc/libsightpy_b200_pt.so python bench.py $S2 2>&1 >/dev/null | grep warp-cycles | tail -4
echo "== launch blocking"; CUDA_LAUNCH_BLOCKING=1 python bench.py $S2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['ms_per_step'], d['config']['level_ms_rank0'])"
for cc in none all; do
echo "== ncu single metric cache-control $cc"
ncu --metrics gpu__time_duration.sum --print-units base --clock-control none --cache-control $cc --csv --log-file gpurun_out/run2_$cc.csv python bench.py $S2 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/run2_$cc.csv")) if len(r)>5]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
print([(r[ki][:22], r[vi]) for r in rows[1:14]])
PY
done
echo "== ncu 5 metrics"
M="gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum"
ncu --metrics $M --print-units base --clock-control none --csv --log-file gpurun_out/run2_5m.csv python bench.py $S2 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/run2_5m.csv")) if len(r)>5]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); mi=h.index("Metric Name")
print([(r[ki][:22], r[vi]) for r in rows[1:] if r[mi]=="gpu__time_duration.sum"][:13])
PY
