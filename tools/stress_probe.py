#!/usr/bin/env python
"""Development aid: render a reduced-resolution stress scene once (for ncu)."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
for p in (REPO, REPO / "python-raytracer_b200", REPO / "tests"):
    sys.path.insert(0, str(p))
import scenes, sightpy
from sightpy.backend import NativeScene
from sightpy.flatten import flatten_scene
w, h, spp = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
nat = NativeScene(flatten_scene(scenes.stress(sightpy, width=w, height=h)))
if len(sys.argv) > 4:
    nat.set_option("bvh", int(sys.argv[4]))          # 0: exhaustive multi-chunk walk
nat.render_samples(0, 1, seed=0)                     # warm-up (queue sizing probe)
st = nat.render_samples(0, spp, seed=0)
print(st["rays_total"], st["device_ms"], st["rays_total"] / st["device_ms"] / 1e3, "Mrays/s", [round(x, 1) for x in st["level_ms"]])
