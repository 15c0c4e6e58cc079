#!/usr/bin/env python
"""Throughput of the five BASELINE.json configurations on one GPU (device-resident scene, frame resolved
on the device).  Writes gpurun_out/configs.json; the headline configuration (4) is what bench.py times.

    python tools/bench_configs.py [--stress-spp N] [--only name ...]
"""
import argparse
import json
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (REPO, REPO / "python-raytracer_b200", REPO / "tests"):
    sys.path.insert(0, str(p))

import scenes  # noqa: E402
import sightpy  # noqa: E402
from sightpy.backend import NativeScene  # noqa: E402
from sightpy.flatten import flatten_scene  # noqa: E402

CONFIGS = [
    # name, builder, kwargs, spp, repeats
    ("1: example1 400x300 1spp", scenes.example1, dict(width=400, height=300), 1, 5),
    ("2a: example2 1920x1080 7spp", scenes.example2, dict(width=1920, height=1080), 7, 3),
    ("2b: example3 1920x1080 4spp", scenes.example3, dict(width=1920, height=1080), 4, 3),
    ("3a: example4 3840x2160 16spp", scenes.example4, dict(width=3840, height=2160), 16, 3),
    ("3b: example3+normalmap 3840x2160 16spp", scenes.example3, dict(width=3840, height=2160, normalmap=True), 16, 3),
    ("4: cornell 1920x1080 256spp", scenes.cornell, dict(width=1920, height=1080), 256, 1),
    ("5: stress 4096 spheres + 2048 triangles 3840x2160", scenes.stress, dict(width=3840, height=2160), None, 1),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stress-spp", type=int, default=2)
    ap.add_argument("--only", nargs="*")
    args = ap.parse_args()
    out = {}
    for name, builder, kw, spp, reps in CONFIGS:
        if args.only and not any(name.startswith(o) for o in args.only):
            continue
        spp = spp or args.stress_spp
        t0 = time.perf_counter()
        scene = builder(sightpy, **kw)
        flat = flatten_scene(scene)
        t_build = time.perf_counter() - t0
        nat = NativeScene(flat)
        nat.render_samples(0, 1, seed=0)                       # warm-up (also sizes the chunks)
        best = None
        for _ in range(reps):
            st = nat.render_samples(0, spp, seed=0)
            nat.resolve_on_device(spp)
            if best is None or st["device_ms"] < best["device_ms"]:
                best = st
        nat.close()
        prim = best["rays_per_depth"][0]
        out[name] = dict(
            colliders=len(flat.colliders), spp=spp, primaries=prim, rays=best["rays_total"],
            rays_per_primary=best["rays_total"] / prim, shadow_rays=best["shadow_rays"],
            s_per_frame=best["device_ms"] / 1e3, mrays_per_s=best["rays_total"] / best["device_ms"] / 1e3,
            ray_collider_tests_per_s=best["rays_total"] * len(flat.colliders) / (best["device_ms"] / 1e3),
            level_ms=best["level_ms"], chunks=best["chunks"], kernel_launches=best["kernel_launches"],
            host_scene_build_s=t_build)
        print(name, json.dumps({k: v for k, v in out[name].items() if k != "level_ms"}), flush=True)
    (REPO / "gpurun_out").mkdir(exist_ok=True)
    (REPO / "gpurun_out" / "configs.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
