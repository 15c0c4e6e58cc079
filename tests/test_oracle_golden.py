"""CPU: the oracle (test infrastructure) replayed against fixtures recorded from the real reference.

The fixtures were produced by tests/golden/make_golden.py, which runs the unmodified reference
(`/root/reference`, only present in the build container) on float32-representable primary rays.
"""
import hashlib
import json

import numpy as np
import pytest
from conftest import GOLDEN, build_scene, load_golden

from oracle.sightpy_oracle import Oracle, tonemap_u8
from sightpy.flatten import flatten_scene

REPORT = json.loads((GOLDEN / "golden_report.json").read_text())
SCENES = [k for k, v in REPORT.items() if isinstance(v, dict)]


@pytest.mark.parametrize("name", SCENES)
def test_oracle_reproduces_reference_radiance(name):
    g = load_golden(name)
    scene = build_scene(name, REPORT[name]["size"])
    np.random.seed(int(g["seed"]))
    out = Oracle(flatten_scene(scene), rng="legacy").trace(g["origins"], g["dirs"])
    assert np.array_equal(out["hit_id"], g["hit_id"].astype(np.int32))
    assert np.array_equal(np.isfinite(out["t"]), np.isfinite(g["t"]))
    fin = np.isfinite(g["t"])
    np.testing.assert_allclose(out["t"][fin], g["t"][fin], rtol=0, atol=1e-12)
    # float64 restatement, same operation order: agreement to rounding noise
    np.testing.assert_allclose(out["rgb"], g["rgb"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("fixture,scene_args", [
    ("camera_example1", None),
    ("camera_lens", dict(look_from=(1.0, 2.0, 3.0), look_at=(0.0, 0.5, -1.0), screen_width=48, screen_height=40,
                         field_of_view=55.0, aperture=0.3, focal_distance=4.0)),
])
def test_oracle_camera_matches_reference(fixture, scene_args):
    import sightpy
    g = load_golden(fixture)
    if scene_args is None:
        scene = build_scene("example1", (64, 48))
    else:
        scene = sightpy.Scene()
        a = dict(scene_args)
        scene.add_Camera(look_from=sightpy.vec3(*a.pop("look_from")), look_at=sightpy.vec3(*a.pop("look_at")), **a)
    np.random.seed(int(g["seed"]))
    O, D, _ = Oracle(flatten_scene(scene), rng="legacy").camera_rays()
    np.testing.assert_allclose(O.T, g["origin"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(D.T, g["dir"], rtol=0, atol=1e-14)


def test_oracle_tonemap_matches_reference():
    g = load_golden("tonemap")
    assert np.array_equal(tonemap_u8(g["linear"], 64, 64).reshape(-1, 3), g["srgb8"])


def test_skybox_blur_matches_reference_digest():
    """Host preprocessing: our cube-map blur is byte-identical to blur_background.py's output."""
    from sightpy.imaging import DECODE_LINEAR, TextureImage, blur_skybox_u8, open_rgb8
    tex = TextureImage(blur_skybox_u8(open_rgb8("sightpy/backgrounds/lake.png"), 10.0, "lake.png"), DECODE_LINEAR)
    arr = np.ascontiguousarray(tex.as_float())
    assert list(arr.shape) == REPORT["blur_lake_shape"]
    assert hashlib.sha256(arr.tobytes()).hexdigest() == REPORT["blur_lake_sha256"]


def test_philox_known_answer():
    """Random123 known-answer vectors for Philox4x32-10 (kat_vectors: zero and all-ones inputs)."""
    from oracle.sightpy_oracle import philox4x32
    out = philox4x32(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in out] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    out = philox4x32(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF)
    assert [int(x) for x in out] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
