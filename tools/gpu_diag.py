#!/usr/bin/env python
"""Development diagnostics (run on the GPU box): per-scene parity numbers of the CUDA path against
the golden fixtures (reference outputs) and the float64 oracle.  Writes gpurun_out/diag.json."""
import json
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
for p in (REPO, REPO / "python-raytracer_b200", REPO / "tests"):
    sys.path.insert(0, str(p))

from conftest import GOLDEN, build_scene, load_golden  # noqa: E402
from oracle.sightpy_oracle import Oracle, tonemap_u8  # noqa: E402
from sightpy.backend import NativeScene, measure_peaks  # noqa: E402
from sightpy.flatten import flatten_scene  # noqa: E402

REPORT = json.loads((GOLDEN / "golden_report.json").read_text())
out = {}
names = sys.argv[1:] or [k for k, v in REPORT.items() if isinstance(v, dict)]
for name in names:
    g = load_golden(name)
    scene = build_scene(name, REPORT[name]["size"])
    flat = flatten_scene(scene)
    nat = NativeScene(flat)
    t0 = time.time()
    res = nat.trace(g["origins"], g["dirs"], seed=5)
    dt = time.time() - t0
    orc = Oracle(flat, rng="philox", seed=5).trace(g["origins"], g["dirs"])
    hit_ref = g["hit_id"].astype(np.int32)
    d_ref = np.abs(res["rgb"].astype(np.float64) - g["rgb"]).max(axis=1)
    d_orc = np.abs(res["rgb"].astype(np.float64) - orc["rgb"]).max(axis=1)
    fin = np.isfinite(g["t"])
    rel_t = np.abs(res["t"][fin] - g["t"][fin]) / np.maximum(g["t"][fin], 1e-9)
    out[name] = dict(
        rays=int(len(hit_ref)), seconds=dt,
        hit_mismatch_vs_reference=int((res["hit_id"] != hit_ref).sum()),
        hit_mismatch_vs_oracle=int((res["hit_id"] != orc["hit_id"]).sum()),
        t_rel_max=float(rel_t.max(initial=0.0)),
        frac_over_1e3_vs_reference=float((d_ref > 1e-3).mean()),
        frac_over_1e3_vs_oracle_philox=float((d_orc > 1e-3).mean()),
        max_vs_oracle=float(d_orc.max()), median_vs_oracle=float(np.median(d_orc)),
        mean_gpu=float(res["rgb"].mean()), mean_oracle=float(orc["rgb"].mean()), mean_reference=float(g["rgb"].mean()),
        stats=res["stats"],
    )
    worst = np.argsort(-d_orc)[:5]
    out[name]["worst"] = [dict(i=int(i), gpu=res["rgb"][i].tolist(), oracle=orc["rgb"][i].tolist(),
                               hit=int(res["hit_id"][i]), hit_o=int(orc["hit_id"][i])) for i in worst]
    print(name, json.dumps({k: v for k, v in out[name].items() if k not in ("stats", "worst")}), flush=True)
    nat.close()
try:
    out["peaks"] = measure_peaks()
    print("peaks", out["peaks"])
except Exception as e:  # noqa: BLE001
    print("peaks failed", e)
(REPO / "gpurun_out").mkdir(exist_ok=True)
(REPO / "gpurun_out" / "diag.json").write_text(json.dumps(out, indent=1))
