#!/usr/bin/env python
"""Frame rate of an animation loop through the public API (update_scene -> Scene.render, animation.py) and the
first-frame latency of example4 (sky-box blur included).  usage: anim_probe.py [frames]"""
import contextlib, io, sys, time
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parent.parent
for p in (REPO, REPO / "python-raytracer_b200", REPO / "tests"):
    sys.path.insert(0, str(p))
import scenes, sightpy

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
quiet = contextlib.redirect_stdout(io.StringIO())
with quiet:
    scenes.example1(sightpy, width=64, height=48).render(samples_per_pixel=1)      # CUDA context, library load
    t0 = time.perf_counter(); sc4 = scenes.example4(sightpy, width=3840, height=2160); t1 = time.perf_counter()
    img = sc4.render(samples_per_pixel=16); t2 = time.perf_counter()
    img = sc4.render(samples_per_pixel=16); t3 = time.perf_counter()
print("example4 4K 16 spp: scene build %.1f ms, first render %.1f ms, second render %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))

for what in ("camera", "sphere"):
    sc = scenes.example1(sightpy, width=400, height=300)
    with quiet:
        sc.render(samples_per_pixel=6)
        t0 = time.perf_counter()
        for i in range(n):
            t = i / n
            if what == "camera":
                sc.camera.look_from = sightpy.vec3(2.5 * np.sin(t), 0.25, 2.5 * np.cos(t) - 1.5)
                sc.add_Camera(sc.camera.look_from, sightpy.vec3(0.0, 0.25, -3.0), screen_width=400, screen_height=300)
            else:
                sc.scene_primitives[0].collider_list[0].center = sightpy.vec3(-0.75, 0.1 + 0.3 * t, -3.0)
            sc.invalidate()
            img = sc.render(samples_per_pixel=6)
        dt = time.perf_counter() - t0
    print("example1 400x300 6 spp, %s moves every frame: %.0f frames/s (%.2f ms per frame, device %.2f ms)" % (what, n / dt, dt / n * 1e3, sc.last_stats["device_ms"]))
