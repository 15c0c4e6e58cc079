#!/usr/bin/env python
"""Debugging aid: compare GPU and oracle radiance truncated after k recursion levels."""
import json, sys
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parent.parent
for p in (REPO, REPO / "python-raytracer_b200", REPO / "tests"):
    sys.path.insert(0, str(p))
from conftest import GOLDEN, build_scene, load_golden
from oracle.sightpy_oracle import Oracle
from sightpy.backend import NativeScene
from sightpy.flatten import flatten_scene
REPORT = json.loads((GOLDEN / "golden_report.json").read_text())
name = sys.argv[1] if len(sys.argv) > 1 else "cornell"
g = load_golden(name)
flat = flatten_scene(build_scene(name, REPORT[name]["size"]))
nat = NativeScene(flat)
for k in range(1, 8):
    nat.set_option("max_levels", k)
    res = nat.trace(g["origins"], g["dirs"], seed=5)
    orc = Oracle(flat, rng="philox", seed=5); orc.max_levels = k
    want = orc.trace(g["origins"], g["dirs"])
    err = np.abs(res["rgb"] - want["rgb"]).max(axis=1)
    print(f"k={k} mean gpu {res['rgb'].mean():.6f} oracle {want['rgb'].mean():.6f} frac>1e-3 {np.mean(err>1e-3):.4f} "
          f"gpu rays {res['stats']['rays_per_depth']} oracle rays {[orc.rays_per_depth[d] for d in sorted(orc.rays_per_depth)]}")
    if np.mean(err > 1e-3) > 0.01:
        for i in np.argsort(-err)[:6]:
            print("   ray", i, "hit", res["hit_id"][i], "gpu", res["rgb"][i], "oracle", want["rgb"][i])
        break
