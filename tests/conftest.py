import os
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
for p in (REPO, REPO / "python-raytracer_b200", REPO / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

GOLDEN = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_golden(name):
    with np.load(GOLDEN / f"{name}.npz") as z:
        return {k: z[k] for k in z.files}


_SCENE_CACHE = {}


def build_scene(name, size):
    """Scene from tests/scenes.py by golden name ('example3_normalmap', 'cornell_mc', ...)."""
    import scenes
    import sightpy
    key = (name, tuple(size))
    if key not in _SCENE_CACHE:
        base, _, variant = name.partition("_")
        kwargs = {}
        if variant == "normalmap":
            kwargs["normalmap"] = True
        if variant == "mc":
            kwargs["mc"] = True
        if variant.isdigit():
            kwargs["seed"] = int(variant)
        _SCENE_CACHE[key] = scenes.BUILDERS[base](sightpy, width=size[0], height=size[1], **kwargs)
    return _SCENE_CACHE[key]
