#!/usr/bin/env python
"""Development aid: where does the end-to-end Scene.render time go?"""
import sys, time
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
for p in (REPO, REPO / "python-raytracer_b200", REPO / "tests"):
    sys.path.insert(0, str(p))
import scenes, sightpy
from sightpy.backend import NativeScene
from sightpy.flatten import flatten_scene
from PIL import Image
scene = scenes.cornell(sightpy, width=1920, height=1080)
for it in range(4):
    t = [time.perf_counter()]
    flat = flatten_scene(scene); t.append(time.perf_counter())
    nat = NativeScene(flat); t.append(time.perf_counter())
    srgb, lin, st = nat.render(16, 0, want_linear=False); t.append(time.perf_counter())
    img = Image.fromarray(srgb, "RGB"); t.append(time.perf_counter())
    nat.close(); t.append(time.perf_counter())
    print(it, "flatten %.1f ms, upload/commit %.1f ms, render+resolve %.1f ms (device %.1f ms, chunks %d), PIL %.1f ms, destroy %.1f ms" % (
        (t[1]-t[0])*1e3, (t[2]-t[1])*1e3, (t[3]-t[2])*1e3, st["device_ms"], st["chunks"], (t[4]-t[3])*1e3, (t[5]-t[4])*1e3))
