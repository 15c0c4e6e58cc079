#!/usr/bin/env python
"""Converged Cornell-box golden image from the REAL reference (build container only).

Renders the scene of tests/scenes.py:cornell at 64x64 with the unmodified reference
(`Camera.get_ray` + `get_raycolor`, numpy global RNG, one seed per sample) in two independent halves of
128 samples each, and stores both half-means.  Their difference is the reference's own Monte-Carlo noise
(split-half RMSE), which is the yardstick of tests/test_gpu_parity.py::test_cornell_matches_reference_converged_image.

Usage: python tests/golden/make_converged.py            (about 2 minutes on 8 cores)
"""
import os
import sys
from multiprocessing import get_context
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
SIZE, SPP_HALF = (64, 64), 128


def one_sample(seed):
    import make_golden as mg
    ref = mg.load_reference()
    cwd = os.getcwd()
    os.chdir(mg.REF_ROOT)
    try:
        scene = mg.build("cornell", ref, SIZE)
    finally:
        os.chdir(cwd)
    np.random.seed(1000 + seed)
    ray = scene.camera.get_ray(scene.n)
    c = ref.get_raycolor(ray, scene)
    n = SIZE[0] * SIZE[1]
    return np.stack([np.broadcast_to(np.asarray(k, dtype=np.float64), (n,)) for k in (c.x, c.y, c.z)])


def main():
    with get_context("spawn").Pool(os.cpu_count()) as pool:
        samples = pool.map(one_sample, range(2 * SPP_HALF))
    a = np.mean(samples[:SPP_HALF], axis=0)
    b = np.mean(samples[SPP_HALF:], axis=0)
    np.savez_compressed(HERE / "cornell_converged_64x64.npz", half_a=a.astype(np.float32), half_b=b.astype(np.float32),
                        spp_half=np.int64(SPP_HALF), width=np.int64(SIZE[0]), height=np.int64(SIZE[1]))
    print("split-half RMSE", float(np.sqrt(np.mean((a - b) ** 2))), "mean", float(((a + b) / 2).mean()))


if __name__ == "__main__":
    main()
