"""Ray bundles and the integrator entry point (reference: sightpy/ray.py:7-163).

``get_raycolor(ray, scene)`` keeps its reference meaning — linear radiance carried by each ray of
a bundle — but is evaluated by the CUDA wavefront path tracer through ``sp_trace``.
"""
import numpy as np

from .vec import vec3, rgb

__all__ = ["Ray", "Hit", "get_raycolor", "get_distances"]


class Ray:
    """SoA bundle of rays plus the batch-level bounce counters (ray.py:7-33)."""

    def __init__(self, origin, dir, depth, n, reflections, transmissions, diffuse_reflections):
        self.length = max(len(origin), len(dir), len(n))
        shape = [self.length]
        self.origin = origin.broadcast_to(shape)
        self.dir = dir.broadcast_to(shape)
        self.n = n.broadcast_to(shape)
        self.depth = depth
        self.reflections = reflections
        self.transmissions = transmissions
        self.diffuse_reflections = diffuse_reflections

    def __len__(self):
        return self.length

    def _like(self, origin, dir, n):
        return Ray(origin, dir, self.depth, n, self.reflections, self.transmissions, self.diffuse_reflections)

    def extract(self, hit_check):
        return self._like(self.origin.extract(hit_check), self.dir.extract(hit_check), self.n.extract(hit_check))

    def __getitem__(self, ind):
        return self._like(self.origin[ind], self.dir[ind], self.n[ind])

    @staticmethod
    def where(cond, x, y):
        if x.depth != y.depth:
            raise ValueError("Both rays must have same depth")
        return Ray(vec3.where(cond, x.origin, y.origin), vec3.where(cond, x.dir, y.dir), x.depth,
                   vec3.where(cond, x.n, y.n), max(x.reflections, y.reflections),
                   max(x.transmissions, y.transmissions), max(x.diffuse_reflections, y.diffuse_reflections))

    @staticmethod
    def concatenate(rays):
        if not all(r.depth == rays[0].depth for r in rays):
            print("All rays must have same depth!")
        return Ray(vec3.concatenate([r.origin for r in rays]), vec3.concatenate([r.dir for r in rays]),
                   rays[0].depth, vec3.concatenate([r.n for r in rays]),
                   max(r.reflections for r in rays), max(r.transmissions for r in rays),
                   max(r.diffuse_reflections for r in rays))


class Hit:
    """Ray/surface intersection record (ray.py:97-119); produced on the GPU, kept for API parity."""

    def __init__(self, distance, orientation, material, collider, surface):
        self.distance, self.orientation = distance, orientation
        self.material, self.collider, self.surface = material, collider, surface
        self.u = self.v = self.N = self.point = None


def _bundle_arrays(ray):
    o = np.stack([np.asarray(c, dtype=np.float32) for c in ray.origin.components()], axis=1)
    d = np.stack([np.asarray(c, dtype=np.float32) for c in ray.dir.components()], axis=1)
    return np.ascontiguousarray(o), np.ascontiguousarray(d)


def get_raycolor(ray, scene, seed=0):
    """Linear radiance of every ray of ``ray`` (camera-level bundles, depth 0) -> vec3 of arrays."""
    if ray.depth != 0 or ray.diffuse_reflections != 0:
        raise NotImplementedError("the GPU integrator traces bundles that start at depth 0")
    o, d = _bundle_arrays(ray)
    out = scene._backend().trace(o, d, seed)["rgb"]
    return rgb(out[:, 0].astype(np.float64), out[:, 1].astype(np.float64), out[:, 2].astype(np.float64))


def get_distances(ray, scene):
    """Grey map of nearest-hit distances clipped at 10 (ray.py:151-163)."""
    o, d = _bundle_arrays(ray)
    t = scene._backend().trace(o, d, 0, want_rgb=False)["t"].astype(np.float64)
    g = np.where(t <= 10, t, 10) / 10
    return rgb(g, g, g)
