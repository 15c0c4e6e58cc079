"""Scene container and render driver (reference: sightpy/scene.py:29-166).

The bookkeeping (``add``, ``add_*Light``, ``add_Background``) is the reference's; ``render`` is the
drop-in boundary: the scene is flattened to POD records, uploaded once through the C ABI, and
the frame comes back as bytes.  Under ``torchrun`` (WORLD_SIZE > 1) every rank renders a
contiguous sample range on its own GPU and the float accumulation buffers are summed with one
NCCL reduce before rank 0 tonemaps (see parallel.py).
"""
import time

import numpy as np
from PIL import Image

from . import lights
from .camera import Camera
from .environment import Panorama, SkyBox
from .ray import get_distances
from .vec import rgb, vec3

__all__ = ["Scene"]


class Scene:
    def __init__(self, ambient_color=rgb(0.01, 0.01, 0.01), n=vec3(1.0, 1.0, 1.0)):
        self.scene_primitives = []
        self.collider_list = []
        self.shadowed_collider_list = []
        self.Light_list = []
        self.importance_sampled_list = []
        self.ambient_color = ambient_color
        self.n = n                      # refractive index of the ambient medium
        # Frame seed of the counter-based RNG.  None (default): every render() of this Scene draws a fresh sample set
        # (seed = number of frames rendered so far), as successive renders of the reference do with numpy's global
        # stream, so averaging several renders reduces noise; an int pins it (the same frame every time).
        self.seed = None
        self.frames_rendered = 0
        self.last_stats = None
        self._native = None
        self._stale = True

    # ---- description ---------------------------------------------------------------------
    def add_Camera(self, look_from, look_at, **kwargs):
        self.camera = Camera(look_from, look_at, **kwargs)
        self.camera._scene = self
        self._dirty()

    def add_PointLight(self, pos, color):
        self.Light_list.append(lights.PointLight(pos, color))
        self._dirty()

    def add_DirectionalLight(self, Ldir, color):
        self.Light_list.append(lights.DirectionalLight(Ldir.normalize(), color))
        self._dirty()

    def add(self, primitive, importance_sampled=False):
        self.scene_primitives.append(primitive)
        self.collider_list.extend(primitive.collider_list)
        if importance_sampled == True:  # noqa: E712  (reference semantics: only literal True-likes)
            self.importance_sampled_list.append(primitive)
        if primitive.shadow == True:  # noqa: E712
            self.shadowed_collider_list.extend(primitive.collider_list)
        self._dirty()

    def add_Background(self, img, light_intensity=0.0, blur=0.0, spherical=False):
        cls = Panorama if spherical else SkyBox
        primitive = cls(img, light_intensity=light_intensity, blur=blur)
        self.scene_primitives.append(primitive)
        self.collider_list.extend(primitive.collider_list)
        self._dirty()

    # ---- backend plumbing ------------------------------------------------------------------
    # The reference deep-copies the scene on every render (scene.py:85), so whatever a script changed in place —
    # primitive.rotate, camera attributes, material fields — is what gets rendered.  Here the committed device copy
    # is kept between renders; scenes of up to AUTO_REFLATTEN colliders are flattened again on every render (tens of
    # microseconds per collider) and only what differs is handed to the device (backend.NativeScene.update); larger
    # scenes are re-described after add*() calls or an explicit invalidate().
    AUTO_REFLATTEN = 512

    def _dirty(self):
        self._stale = True

    def invalidate(self, full=False):
        """Call after mutating primitives in place (e.g. in an animation's update_scene).  full=True also drops the
        device copy, so that the next render describes and uploads the whole scene again."""
        self._stale = True
        if full and self._native is not None:
            self._native.close()
            self._native = None

    def _backend(self):
        from .backend import NativeGroup, NativeScene, configured_devices
        from .flatten import flatten_scene
        if self._native is None:
            devices = configured_devices()
            flat = flatten_scene(self)
            self._native = NativeGroup(flat, devices) if len(devices) > 1 else NativeScene(flat, device=devices[0])
        elif self._stale or len(self.collider_list) <= self.AUTO_REFLATTEN:
            self._native.update(flatten_scene(self))
        self._stale = False
        return self._native

    def _frame_seed(self):
        return self.frames_rendered if self.seed is None else int(self.seed)

    # ---- rendering -------------------------------------------------------------------------
    def render(self, samples_per_pixel, progress_bar=False, batch_size=None):
        """Render ``samples_per_pixel`` jittered samples per pixel -> PIL "RGB" image.
        ``progress_bar`` / ``batch_size`` are accepted for signature compatibility; the GPU
        backend chunks the wavefront itself and any sample count is valid."""
        print("Rendering...")
        t0 = time.time()
        from .parallel import render_frame
        srgb8, stats = render_frame(self._backend(), int(samples_per_pixel), self._frame_seed())
        self.frames_rendered += 1
        self.last_stats = stats
        print("Render Took", time.time() - t0)
        return Image.fromarray(srgb8, "RGB")

    def get_aovs(self, sample=0):
        """Debug buffers of one primary ray per pixel: dict(hit_id H x W int32 (index into collider_list, -1 = none),
        t H x W float32 (inf = none), normal H x W x 3 float32 (collider normal facing the ray))."""
        return self._backend().aovs(sample, self._frame_seed())

    def get_distances(self):
        """Debug depth map (scene.py:142-166)."""
        print("Rendering...")
        t0 = time.time()
        t = self._backend().distances(self._frame_seed()).astype(np.float64)
        g = np.where(t <= 10, t, 10) / 10
        print("Render Took", time.time() - t0)
        h, w = self.camera.screen_height, self.camera.screen_width
        plane = (255 * np.clip(g, 0, 1).reshape(h, w)).astype(np.uint8)
        return Image.merge("RGB", [Image.fromarray(plane, "L")] * 3)
