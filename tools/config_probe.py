#!/usr/bin/env python
"""Development aid: render one of tests/scenes.py once (for ncu).  usage: config_probe.py example4 1920 1080 2"""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
for p in (REPO, REPO / "python-raytracer_b200", REPO / "tests"):
    sys.path.insert(0, str(p))
import scenes, sightpy
from sightpy.backend import NativeScene
from sightpy.flatten import flatten_scene
name, w, h, spp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
nat = NativeScene(flatten_scene(getattr(scenes, name)(sightpy, width=w, height=h)))
nat.render_samples(0, 1, seed=0)
st = nat.render_samples(0, spp, seed=0)
print(st["rays_total"], st["device_ms"], st["rays_total"] / st["device_ms"] / 1e3, "Mrays/s", [round(x, 2) for x in st["level_ms"]], st["chunks"])
