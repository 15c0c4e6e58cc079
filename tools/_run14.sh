timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 300 2>&1 | tail -8
timeout 500 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2b_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], d["e2e"]["value"], d["clocks"])
for k,v in d["configs"].items(): print(k, round(v["Mrays_per_s"]), v["s_per_frame"])
PY
