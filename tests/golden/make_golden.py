#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the REAL reference implementation.

Runs only in the build container, where lmondada/Python-Raytracer is mounted read-only at
/root/reference (it does not exist on the GPU box, which is why the outputs are committed).

For every scene in tests/scenes.py the script
  1. builds the scene twice from the same builder: with the reference's classes (imported under
     the alias ``refsightpy`` so it cannot be confused with this repo's ``sightpy``) and with ours;
  2. draws one set of jittered primary rays, rounds them to float32 (both renderers then see
     bit-identical, float32-representable rays: SURVEY §7 "parity protocol"); the float64 side
     renormalises the directions, as the reference's own camera rays are unit length to 1e-16;
  3. evaluates the reference's ``get_raycolor`` on them under ``np.random.seed(SEED)`` and records
     linear radiance, nearest collider index and hit distance per ray;
  4. evaluates the oracle (rng="legacy", same seed) on the flattening of BOTH scene objects and
     reports the largest deviation from the reference — the oracle's pin.
Nothing here modifies /root/reference; the only compatibility shim is numpy-2's np.abs(vec3)
(SURVEY App. C).

Usage:  python tests/golden/make_golden.py [scene ...]
"""
import hashlib
import importlib.util
import json
import os
import sys
from functools import reduce
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF_ROOT = Path("/root/reference")
SEED = 7
sys.path.insert(0, str(REPO / "python-raytracer_b200"))
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))

GOLDEN_SIZES = {  # (width, height) of the fixture renders
    "example1": (128, 96), "example2": (128, 96), "example3": (128, 96), "example4": (128, 96),
    "example3_normalmap": (96, 72), "example2_mc": (96, 72), "triangles": (96, 72),
    "cornell": (40, 40), "cornell_mc": (32, 32),
}
# seeded random scenes (tests/scenes.py: fuzz): every collider type / deterministic material at random poses
N_FUZZ = 12
GOLDEN_SIZES.update({f"fuzz_{i}": (48, 36) for i in range(N_FUZZ)})
N_FUZZ_MC = 8
GOLDEN_SIZES.update({f"fuzzmc_{i}": (32, 24) for i in range(N_FUZZ_MC)})


def load_reference():
    spec = importlib.util.spec_from_file_location(
        "refsightpy", REF_ROOT / "sightpy" / "__init__.py",
        submodule_search_locations=[str(REF_ROOT / "sightpy")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["refsightpy"] = mod
    cwd = os.getcwd()
    os.chdir(REF_ROOT)              # asset paths of the reference are CWD-relative
    sys.dont_write_bytecode = True
    try:
        spec.loader.exec_module(mod)
    finally:
        os.chdir(cwd)
    vec3 = mod.vec3

    def array_ufunc(self, ufunc, method, *inputs, **kw):   # numpy>=2 shim, SURVEY App. C
        if ufunc is np.absolute and method == "__call__":
            return abs(self)
        return NotImplemented
    vec3.__array_ufunc__ = array_ufunc
    return mod


def build(name, ns, size):
    import scenes
    w, h = size
    base, _, variant = name.partition("_")
    kwargs = {}
    if variant == "normalmap":
        kwargs["normalmap"] = True
    if variant == "mc":
        kwargs["mc"] = True
    if variant.isdigit():
        kwargs["seed"] = int(variant)
    return scenes.BUILDERS[base](ns, width=w, height=h, **kwargs)


def reference_trace(ref, scene, O32, D32):
    """get_raycolor + per-collider intersect of the reference on float32-representable rays."""
    v = ref.vec3
    O = v(*(O32[:, k].astype(np.float64) for k in range(3)))
    # float32-representable directions, renormalised in float64: the reference's hit points are
    # O + D*|D t|, so |D| must be 1 to rounding as it is for Camera.get_ray's own rays (see
    # Oracle.trace, which applies the same renormalisation)
    D64 = D32.astype(np.float64)
    D64 = D64 / np.sqrt((D64 * D64).sum(axis=1, keepdims=True))
    D = v(*(D64[:, k] for k in range(3)))
    ray = ref.Ray(O, D, 0, scene.n, 0, 0, 0)
    dists = [c.intersect(ray.origin, ray.dir)[0] for c in scene.collider_list]
    nearest = reduce(np.minimum, dists)
    hit = np.full(len(nearest), -1, dtype=np.int32)
    for ci in reversed(range(len(dists))):
        hit[(nearest != ref.FARAWAY) & (dists[ci] == nearest)] = ci
    np.random.seed(SEED)
    c = ref.get_raycolor(ray, scene)
    rgb = np.stack([np.broadcast_to(np.asarray(k, dtype=np.float64), nearest.shape) for k in (c.x, c.y, c.z)], axis=1)
    return rgb, hit, np.where(nearest == ref.FARAWAY, np.inf, nearest)


def main(argv):
    import sightpy as ours
    from sightpy.flatten import flatten_scene
    from oracle.sightpy_oracle import Oracle, tonemap_u8

    ref = load_reference()
    names = argv or list(GOLDEN_SIZES)
    report = json.loads((HERE / "golden_report.json").read_text()) if argv else {}
    for name in names:
        size = GOLDEN_SIZES[name]
        cwd = os.getcwd()
        os.chdir(REF_ROOT)
        try:
            ref_scene = build(name, ref, size)
        finally:
            os.chdir(cwd)
        our_scene = build(name, ours, size)
        flat_ref, flat_ours = flatten_scene(ref_scene), flatten_scene(our_scene)

        # shared, float32-representable primary rays from the reference camera
        np.random.seed(SEED + 1)
        ray = ref_scene.camera.get_ray(ref_scene.n)
        O32 = np.stack([np.broadcast_to(c, (len(ray),)) for c in ray.origin.components()], 1).astype(np.float32)
        D32 = np.stack([c for c in ray.dir.components()], 1).astype(np.float32)

        rgb, hit, t = reference_trace(ref, ref_scene, O32, D32)
        devs = {}
        for label, flat in (("flatten(reference objects)", flat_ref), ("flatten(our objects)", flat_ours)):
            np.random.seed(SEED)
            out = Oracle(flat, rng="legacy").trace(O32, D32)
            finite = np.isfinite(rgb).all(axis=1) & np.isfinite(out["rgb"]).all(axis=1)
            d = np.abs(out["rgb"] - rgb)[finite]
            devs[label] = dict(max_abs=float(d.max()), n_over_1e9=int((d.max(axis=1) > 1e-9).sum()),
                               hit_mismatch=int((out["hit_id"] != hit).sum()),
                               t_max_abs=float(np.abs((out["t"] - t)[np.isfinite(t)]).max(initial=0.0)),
                               t_inf_mismatch=int((np.isfinite(t) != np.isfinite(out["t"])).sum()),
                               nonfinite=int((~finite).sum()))
        report[name] = dict(size=size, rays=int(len(hit)), deviation=devs)
        print(name, json.dumps(report[name]["deviation"], indent=1))
        np.savez_compressed(HERE / f"{name}.npz", origins=O32, dirs=D32, rgb=rgb, hit_id=hit.astype(np.int16),
                            t=t, seed=np.int64(SEED))

    if not argv:
        # camera: reference Camera.get_ray vs oracle camera_rays on the same numpy stream
        os.chdir(REF_ROOT)
        try:
            cam_scene = build("example1", ref, (64, 48))
        finally:
            os.chdir(REPO)
        np.random.seed(SEED)
        r = cam_scene.camera.get_ray(cam_scene.n)
        np.savez_compressed(HERE / "camera_example1.npz",
                            origin=np.stack([np.broadcast_to(c, (len(r),)) for c in r.origin.components()], 1),
                            dir=np.stack(list(r.dir.components()), 1), seed=np.int64(SEED))
        # thin-lens variant
        lens_scene = ref.Scene()
        lens_scene.add_Camera(look_from=ref.vec3(1.0, 2.0, 3.0), look_at=ref.vec3(0.0, 0.5, -1.0), screen_width=48,
                              screen_height=40, field_of_view=55.0, aperture=0.3, focal_distance=4.0)
        np.random.seed(SEED)
        r = lens_scene.camera.get_ray(lens_scene.n)
        np.savez_compressed(HERE / "camera_lens.npz", origin=np.stack(list(r.origin.components()), 1),
                            dir=np.stack(list(r.dir.components()), 1), seed=np.int64(SEED))

        # tonemap: reference sRGB_linear_to_sRGB + uint8 truncation (scene.py:118-140)
        rs = np.random.RandomState(3)
        lin = np.abs(rs.standard_normal((3, 4096))) * rs.choice([1e-3, 0.05, 0.5, 2.0, 20.0], size=4096)
        lin[:, :8] = np.array([[0, 0.00304, 0.0031, 1.0, 0.999, 1.0001, 0.5, 1e-9]] * 3)
        enc = ref.sRGB_linear_to_sRGB(lin)
        u8 = np.stack([(255 * np.clip(c, 0, 1)).astype(np.uint8) for c in enc], axis=1)
        np.savez_compressed(HERE / "tonemap.npz", linear=lin, srgb8=u8)
        assert np.array_equal(tonemap_u8(lin, 64, 64).reshape(-1, 3), u8), "oracle tonemap != reference"

        # skybox blur: digest of the reference's blurred, linearised cube map (lake.png, blur 10)
        from refsightpy.backgrounds.util.blur_background import blur_skybox
        cwd = os.getcwd()
        os.chdir(REF_ROOT)
        try:
            blurred = blur_skybox(ref.load_image("sightpy/backgrounds/lake.png"), 10.0, "lake.png")
        finally:
            os.chdir(cwd)
        report["blur_lake_sha256"] = hashlib.sha256(np.ascontiguousarray(blurred).tobytes()).hexdigest()
        report["blur_lake_shape"] = list(blurred.shape)

        # the reference's own rendered examples (end-to-end acceptance band)
        for i in range(1, 5):
            (HERE / f"EXAMPLE{i}.png").write_bytes((REF_ROOT / "images" / f"EXAMPLE{i}.png").read_bytes())
    (HERE / "golden_report.json").write_text(json.dumps(report, indent=1))
    print("done")


if __name__ == "__main__":
    main(sys.argv[1:])
