"""Grazing / texel-tie mask of the parity protocol (SURVEY §7 "hard parts", item ii).

Nearest-neighbour textures, checker floors with thousands of texels across and silhouette edges turn a 1e-7 relative
change of a hit point into a different texel or a different collider.  Whether a given test ray sits on such a
discontinuity is a property of the *reference*, not of the implementation under test: a ray is masked iff the
reference's own answer for it moves by more than the tolerance — or its nearest collider changes — when the ray's
direction is perturbed by `ulps` float32 ulps (each axis, both signs).  The float64 oracle (pinned to the reference
at 1e-16, tests/test_oracle_golden.py) evaluates the perturbed rays.  Tests then demand *every unmasked ray* within
tolerance and report the masked fraction, instead of granting a flat budget of outliers.
"""
import numpy as np

from oracle.sightpy_oracle import Oracle


def perturbed_directions(D32, ulps):
    """The 6 neighbours of each float32 direction: +-ulps steps along each axis."""
    D32 = np.ascontiguousarray(D32, dtype=np.float32)
    for axis in range(3):
        for sign in (1.0, -1.0):
            Dp = D32.copy()
            target = np.full(len(Dp), np.float32(sign * np.inf), dtype=np.float32)
            for _ in range(ulps):
                Dp[:, axis] = np.nextafter(Dp[:, axis], target)
            yield Dp


def sensitivity_mask(flat, O32, D32, tol=1e-3, ulps=2, seed=0):
    """-> (mask, base): mask[i] is True where the oracle's radiance / hit id of ray i is not stable under the
    perturbation; base is the oracle's answer for the unperturbed rays."""
    base = Oracle(flat, rng="philox", seed=seed).trace(O32, D32)
    mask = np.zeros(len(O32), dtype=bool)
    for Dp in perturbed_directions(D32, ulps):
        out = Oracle(flat, rng="philox", seed=seed).trace(O32, Dp)
        moved = np.abs(out["rgb"] - base["rgb"]).max(axis=1) > tol
        mask |= moved | (out["hit_id"] != base["hit_id"]) | ~np.isfinite(out["rgb"]).all(axis=1)
    return mask, base
