"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, agrees with the binding's struct layouts, and refuses to run without a GPU."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
HEADER = (REPO / "include" / "sightpy_b200.h").read_text()


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as entry
    entry.build()
    from sightpy.backend import load_library
    return load_library()


def declared_functions():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(sp_[a-z0-9_]+)\s*\(", body)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ("sp_init", "sp_scene_create", "sp_scene_commit", "sp_render", "sp_render_samples", "sp_resolve",
                 "sp_trace", "sp_camera_rays", "sp_distances", "sp_last_error", "sp_scene_set_stream"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, f"declared in include/sightpy_b200.h but not exported: {missing}"


def test_struct_layouts_match_binding(lib):
    from sightpy.backend import Stats
    from sightpy.flatten import CAMERA_DT, COLLIDER_DT, LIGHT_DT, MATERIAL_DT, PRIMITIVE_DT
    sizes = (C.c_int32 * 6)()
    assert lib.sp_abi_sizes(sizes) == 0
    assert list(sizes) == [CAMERA_DT.itemsize, MATERIAL_DT.itemsize, PRIMITIVE_DT.itemsize, COLLIDER_DT.itemsize,
                           LIGHT_DT.itemsize, C.sizeof(Stats)]
    assert lib.sp_abi_version() == int(re.search(r"#define SP_ABI_VERSION (\d+)", HEADER).group(1))


def test_no_cpu_fallback(lib):
    """Without a CUDA device sp_init must fail loudly (and scene creation must refuse to proceed)."""
    if lib.sp_device_count() > 0:
        pytest.skip("a CUDA device is visible: the failure path cannot be exercised here")
    assert lib.sp_init(0) != 0
    assert b"no CPU fallback" in lib.sp_last_error()
    from sightpy.backend import NativeScene
    from conftest import build_scene
    from sightpy.flatten import flatten_scene
    with pytest.raises(RuntimeError, match="no CUDA device"):
        NativeScene(flatten_scene(build_scene("cornell", (8, 8))))


def test_package_never_imports_the_oracle():
    pkg = REPO / "python-raytracer_b200"
    offenders = [str(p) for p in pkg.rglob("*.py") if re.search(r"^\s*(from|import)\s+oracle\b", p.read_text(), re.M)]
    offenders += [str(p) for p in (pkg / "csrc").glob("*") if p.suffix in (".cu", ".cuh", ".h")
                  and re.search(r"#include\s+[<\"][^>\"]*oracle", p.read_text())]
    assert not offenders


def test_flatten_cornell_records():
    from conftest import build_scene
    from sightpy.flatten import COLLIDER_CUBOID, COLLIDER_PLANE, COLLIDER_SPHERE, MAT_DIFFUSE, flatten_scene
    flat = flatten_scene(build_scene("cornell", (16, 16)))
    assert [int(t) for t in flat.colliders["type"]] == [COLLIDER_PLANE] * 6 + [COLLIDER_CUBOID, COLLIDER_SPHERE]
    assert len(flat.importance) == 2 and len(flat.media) == 2
    assert flat.max_depth_bound() >= 6
    dif = flat.materials[flat.materials["kind"] == MAT_DIFFUSE]
    assert set(dif["diffuse_rays"]) == {20} and set(dif["max_diffuse_reflections"]) == {2}
    light = flat.primitives[int(flat.importance[0])]
    np.testing.assert_allclose(light["bounded_sphere_radius"], np.hypot(65.0, 52.5))


def test_python_material_subclass_is_rejected():
    import sightpy
    from sightpy.flatten import flatten_scene

    class Custom(sightpy.Material):
        pass

    scene = sightpy.Scene()
    scene.add_Camera(look_from=sightpy.vec3(0, 0, 1), look_at=sightpy.vec3(0, 0, 0), screen_width=8, screen_height=8)
    scene.add(sightpy.Sphere(material=Custom(), center=sightpy.vec3(0, 0, -1), radius=0.5))
    with pytest.raises(TypeError, match="not supported"):
        flatten_scene(scene)
