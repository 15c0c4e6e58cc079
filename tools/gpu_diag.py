#!/usr/bin/env python
"""Development diagnostics (run on the GPU box): per-scene parity numbers of the CUDA path against the golden
fixtures (reference outputs) and the float64 oracle, with the tie mask of tests/tiemask.py at several perturbation
sizes.  Writes gpurun_out/diag.json; the thresholds of tests/test_gpu_parity.py are set from this file
(profiles/r2_parity_measured.json is a committed copy)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
for p in (REPO, REPO / "python-raytracer_b200", REPO / "tests"):
    sys.path.insert(0, str(p))

from conftest import GOLDEN, build_scene, load_golden  # noqa: E402
from oracle.sightpy_oracle import Oracle  # noqa: E402
from sightpy.backend import NativeScene  # noqa: E402
from sightpy.flatten import flatten_scene  # noqa: E402
from tiemask import sensitivity_mask  # noqa: E402

REPORT = json.loads((GOLDEN / "golden_report.json").read_text())
out = {}
names = sys.argv[1:] or [k for k, v in REPORT.items() if isinstance(v, dict)]
for name in names:
    g = load_golden(name)
    scene = build_scene(name, REPORT[name]["size"])
    flat = flatten_scene(scene)
    nat = NativeScene(flat)
    mc = name.startswith(("fuzzmc", "cornell")) or name.endswith("_mc")
    seed = 11 if mc else 0
    t0 = time.time()
    res = nat.trace(g["origins"], g["dirs"], seed=seed)
    dt = time.time() - t0
    orc = Oracle(flat, rng="philox", seed=seed).trace(g["origins"], g["dirs"])
    hit_ref = g["hit_id"].astype(np.int32)
    d_ref = np.abs(res["rgb"].astype(np.float64) - g["rgb"]).max(axis=1)
    d_orc = np.abs(res["rgb"].astype(np.float64) - orc["rgb"]).max(axis=1)
    scale = 1.0 + np.abs(orc["rgb"]).max(axis=1)
    fin = np.isfinite(g["t"])
    rel_t = np.abs(res["t"][fin] - g["t"][fin]) / np.maximum(g["t"][fin], 1e-9)
    row = dict(
        rays=int(len(hit_ref)), seconds=dt, monte_carlo=bool(mc),
        hit_mismatch_vs_reference=int((res["hit_id"] != hit_ref).sum()),
        hit_mismatch_vs_oracle=int((res["hit_id"] != orc["hit_id"]).sum()),
        t_rel_max=float(rel_t.max(initial=0.0)),
        n_over_1e3_vs_reference=int((d_ref > 1e-3).sum()),
        n_over_1e3_vs_oracle_philox=int((d_orc > 1e-3).sum()),
        n_over_1e3_scaled_vs_oracle_philox=int((d_orc > 1e-3 * scale).sum()),
        max_vs_oracle=float(d_orc.max()), median_vs_oracle=float(np.median(d_orc)),
        mean_gpu=float(res["rgb"].mean()), mean_oracle=float(orc["rgb"].mean()), mean_reference=float(g["rgb"].mean()),
    )
    if not mc:
        viol = d_ref > 1e-3
        for ulps in (1, 2, 4):
            mask, _ = sensitivity_mask(flat, g["origins"], g["dirs"], tol=1e-3, ulps=ulps, seed=seed)
            row[f"masked_{ulps}ulp"] = int(mask.sum())
            row[f"unmasked_violations_{ulps}ulp"] = int((viol & ~mask).sum())
    out[name] = row
    print(name, json.dumps(row), flush=True)
    nat.close()
(REPO / "gpurun_out").mkdir(exist_ok=True)
(REPO / "gpurun_out" / "diag.json").write_text(json.dumps(out, indent=1))
