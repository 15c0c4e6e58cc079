"""The UNMODIFIED reference (lmondada/Python-Raytracer) as a timed CPU arm.  TEST / BENCH INFRASTRUCTURE ONLY:
imported by tests/, tests/golden/*.py and bench.py's CPU legs, never by the product path.

The reference is pure Python, so "installing" it is a copy: ``install()`` (called by ``__graft_entry__.build()``
in the build container, where /root/reference is mounted) copies its ``sightpy`` package into the git-ignored
``baseline/_ref/`` — not part of this repository's history, but part of the snapshot ``gpurun`` ships to the GPU box, so
the real reference can be timed on that box's host cores next to the CUDA path.  Nothing in it is edited; the only
compatibility shim is numpy-2's ``np.abs(vec3)``, applied at import time by monkey-patching.

Two measurements:
  * ``trace_rate``: ``get_raycolor`` (sightpy/ray.py:122-148) on the scene's own camera rays, one process per core,
    one sample per process and step; rays = sum of len(ray) over all get_raycolor calls (the metric's definition);
  * ``render_as_shipped``: ``Scene.render`` (sightpy/scene.py:71-140) as the example scripts call it, process pool,
    deep copies, pickling and all.
"""
import importlib.util
import os
import shutil
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
REF_SRC = Path("/root/reference")
REF_DST = REPO / "baseline" / "_ref"
_MODULE = None
_RAYS = [0]


def install():
    """Copy the reference package into baseline/_ref (idempotent).  Returns True when the copy exists afterwards."""
    if (REF_SRC / "sightpy").is_dir():
        dst = REF_DST / "sightpy"
        if not dst.is_dir() or not (dst / "__init__.py").exists():
            REF_DST.mkdir(parents=True, exist_ok=True)
            shutil.copytree(REF_SRC / "sightpy", dst, dirs_exist_ok=True,
                            ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return (REF_DST / "sightpy" / "__init__.py").exists()


def available():
    return (REF_DST / "sightpy" / "__init__.py").exists() or (REF_SRC / "sightpy" / "__init__.py").exists()


def root():
    return REF_DST if (REF_DST / "sightpy" / "__init__.py").exists() else REF_SRC


class in_root:
    """The reference loads its assets with CWD-relative paths (sightpy/textures/texture.py:29)."""

    def __enter__(self):
        self.cwd = os.getcwd()
        os.chdir(root())

    def __exit__(self, *exc):
        os.chdir(self.cwd)


def load():
    """Import the reference under the alias ``refsightpy`` (so it cannot be confused with this repo's ``sightpy``),
    with the numpy-2 shim and a ray counter around get_raycolor."""
    global _MODULE
    if _MODULE is not None:
        return _MODULE
    base = root()
    spec = importlib.util.spec_from_file_location("refsightpy", base / "sightpy" / "__init__.py",
                                                  submodule_search_locations=[str(base / "sightpy")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["refsightpy"] = mod
    sys.dont_write_bytecode = True
    with in_root():
        spec.loader.exec_module(mod)

    def array_ufunc(self, ufunc, method, *inputs, **kw):   # numpy >= 2: np.abs(vec3), SURVEY App. C
        if ufunc is np.absolute and method == "__call__":
            return abs(self)
        return NotImplemented
    mod.vec3.__array_ufunc__ = array_ufunc

    # count rays the way the metric defines them: len(ray) of every get_raycolor call.  The materials import the
    # function by name, so the counting wrapper is installed in every module that holds a reference to it.
    inner = sys.modules["refsightpy.ray"].get_raycolor

    def counted(ray, scene):
        _RAYS[0] += len(ray)
        return inner(ray, scene)
    for name, m in list(sys.modules.items()):
        if name.startswith("refsightpy") and getattr(m, "get_raycolor", None) is inner:
            m.get_raycolor = counted
    _MODULE = mod
    return mod


def build_scene(builder_name, **kw):
    """tests/scenes.py builder with the reference's classes."""
    for p in (REPO / "tests",):
        if str(p) not in sys.path:
            sys.path.insert(0, str(p))
    import scenes
    ref = load()
    with in_root():
        return scenes.BUILDERS[builder_name](ref, **kw)


def _trace_worker(job):
    builder_name, kw, seed = job
    ref = load()
    scene = build_scene(builder_name, **kw)
    np.random.seed(seed)
    _RAYS[0] = 0
    t0 = time.perf_counter()
    ray = scene.camera.get_ray(scene.n)
    ref.get_raycolor(ray, scene)
    return _RAYS[0], time.perf_counter() - t0


def trace_rate(builder_name, kw, n_samples, processes):
    """get_raycolor on n_samples independent samples of the scene's frame, `processes` at a time.
    Returns (rays, wall seconds)."""
    jobs = [(builder_name, kw, 1000 + s) for s in range(n_samples)]
    if processes == 1:
        t0 = time.perf_counter()
        res = [_trace_worker(j) for j in jobs]
        return sum(r for r, _ in res), time.perf_counter() - t0
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(processes) as pool:
        small = dict(kw, width=max(kw.get("width", 64) // 8, 8), height=max(kw.get("height", 64) // 8, 8))
        pool.map(_trace_worker, [(builder_name, small, 0)] * processes)     # spin-up (imports) outside the timing
        t0 = time.perf_counter()
        res = pool.map(_trace_worker, jobs)
        dt = time.perf_counter() - t0
    return sum(r for r, _ in res), dt


def render_as_shipped(builder_name, kw, spp):
    """Scene.render of the reference exactly as an example script runs it.  Returns wall seconds.
    (spp must be a multiple of ceil(spp / cpu_count()): the shipped batching crashes on a ragged last batch.)"""
    import contextlib
    import io
    scene = build_scene(builder_name, **kw)
    np.random.seed(0)
    with in_root(), contextlib.redirect_stdout(io.StringIO()):
        t0 = time.perf_counter()
        scene.render(samples_per_pixel=spp)
        return time.perf_counter() - t0
