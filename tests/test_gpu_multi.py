"""Several GPUs of one node driven by the library itself (sp_init_devices / sp_render_group): needs >= 2 devices, so
the single-GPU runs of the suite skip it; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu` runs it."""
import numpy as np
import pytest

from sightpy.flatten import flatten_scene

pytestmark = pytest.mark.gpu


def _devices():
    from sightpy.backend import visible_devices
    return visible_devices()


@pytest.mark.parametrize("shard", ["samples", "tiles"])
def test_group_frame_equals_single_device_frame(shard):
    """One frame spread over the node's GPUs by sp_render_group (host thread per device, peer-access gather) equals
    the frame one GPU renders alone: same Philox keys per (pixel, sample), so equal up to float addition order."""
    devices = _devices()
    if len(devices) < 2:
        pytest.skip("needs at least two GPUs")
    import scenes
    import sightpy
    from sightpy.backend import NativeGroup, NativeScene
    flat = flatten_scene(scenes.cornell(sightpy, width=200, height=150))
    single = NativeScene(flat, device=devices[0])
    _, want, st1 = single.render(6, seed=3)
    single.close()
    group = NativeGroup(flat, devices)
    group.GROUP_MIN_PRIMARIES = 0
    srgb, got, st = group.render(6, seed=3, shard=shard)
    assert st["devices"] == devices and st["rays_total"] == st1["rays_total"]
    assert st["rays_per_depth"] == st1["rays_per_depth"]
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-6)
    # fewer samples than devices: tiles whatever was asked for
    _, got1, _ = group.render(1, seed=3, shard="samples")
    single = NativeScene(flat, device=devices[-1])              # and a scene on another device alone
    _, want1, _ = single.render(1, seed=3)
    single.close()
    group.close()
    np.testing.assert_allclose(got1, want1, rtol=1e-4, atol=1e-6)


def test_scene_render_uses_every_gpu_without_a_launcher(monkeypatch):
    """`python example.py` on a multi-GPU node: Scene.render spreads a large enough frame over all visible devices
    (SIGHTPY_DEVICES unset, no torchrun), like the reference's render() uses every core (scene.py:80)."""
    devices = _devices()
    if len(devices) < 2:
        pytest.skip("needs at least two GPUs")
    import scenes
    import sightpy
    from sightpy import backend
    monkeypatch.delenv("SIGHTPY_DEVICE", raising=False)
    monkeypatch.delenv("SIGHTPY_DEVICES", raising=False)
    monkeypatch.setattr(backend.NativeGroup, "GROUP_MIN_PRIMARIES", 0)
    sc = scenes.cornell(sightpy, width=160, height=120)
    sc.seed = 2
    img = np.asarray(sc.render(4))
    assert sc.last_stats["devices"] == devices
    monkeypatch.setenv("SIGHTPY_DEVICES", "0")
    one = scenes.cornell(sightpy, width=160, height=120)
    one.seed = 2
    ref = np.asarray(one.render(4))
    assert "devices" not in one.last_stats
    assert np.abs(img.astype(int) - ref.astype(int)).max() <= 1
