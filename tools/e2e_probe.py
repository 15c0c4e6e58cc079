#!/usr/bin/env python
"""Development aid: where does the end-to-end Scene.render time go?  usage: e2e_probe.py [scene] [w] [h] [spp]"""
import sys, time
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
for p in (REPO, REPO / "python-raytracer_b200", REPO / "tests"):
    sys.path.insert(0, str(p))
import scenes, sightpy
from sightpy.backend import NativeScene
from sightpy.flatten import flatten_scene
from PIL import Image
name = sys.argv[1] if len(sys.argv) > 1 else "cornell"
w, h, spp = (int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1920, 1080, 16)
scene = getattr(scenes, name)(sightpy, width=w, height=h)
for it in range(4):
    t = [time.perf_counter()]
    flat = flatten_scene(scene); t.append(time.perf_counter())
    nat = NativeScene(flat); t.append(time.perf_counter())
    srgb, lin, st = nat.render(spp, 0, want_linear=False); t.append(time.perf_counter())
    img = Image.fromarray(srgb, "RGB"); t.append(time.perf_counter())
    nat.close(); t.append(time.perf_counter())
    print(it, "flatten %.1f ms, upload/commit %.1f ms, render+resolve %.1f ms (device %.1f ms, chunks %d), PIL %.1f ms, destroy %.1f ms" % (
        (t[1]-t[0])*1e3, (t[2]-t[1])*1e3, (t[3]-t[2])*1e3, st["device_ms"], st["chunks"], (t[4]-t[3])*1e3, (t[5]-t[4])*1e3))
