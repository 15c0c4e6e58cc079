#!/usr/bin/env python
"""Small all-features workload for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
for p in (REPO, REPO / "python-raytracer_b200", REPO / "tests"):
    sys.path.insert(0, str(p))
import numpy as np
import scenes, sightpy
from sightpy.backend import NativeScene
from sightpy.flatten import flatten_scene
for name, kw in (("cornell", dict(width=32, height=24)), ("example1", dict(width=32, height=24)),
                 ("example4", dict(width=32, height=24)), ("triangles", dict(width=32, height=24)),
                 ("stress", dict(width=16, height=12, n_spheres=400, n_triangles=160, n_collections=2))):
    nat = NativeScene(flatten_scene(getattr(scenes, name)(sightpy, **kw)))
    srgb, lin, st = nat.render(2, seed=1)
    o, d = nat.camera_rays(0, 1)
    tr = nat.trace(o, d, seed=1)
    nat.distances(1)
    nat.close()
    print(name, st["rays_total"], float(lin.mean()), int(tr["hit_id"].max()))
print("done")
