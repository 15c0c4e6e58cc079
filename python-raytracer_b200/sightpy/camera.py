"""Pin-hole / thin-lens camera description (reference: sightpy/camera.py:8-85).

``__init__`` derives the same basis and image-plane extents as the reference.  Primary rays are
generated on the GPU (csrc: camera_ray()); ``get_ray`` fetches one sample's worth of them through
the C ABI so caller code that wants explicit ray bundles keeps working.
"""
import numpy as np

from .vec import vec3

__all__ = ["Camera"]


class Camera:
    def __init__(self, look_from, look_at, screen_width=400, screen_height=300,
                 field_of_view=90.0, aperture=0.0, focal_distance=1.0):
        self.screen_width = screen_width
        self.screen_height = screen_height
        self.aspect_ratio = float(screen_width) / screen_height
        self.look_from, self.look_at = look_from, look_at
        self.camera_width = np.tan(field_of_view * np.pi / 180 / 2.0) * 2.0
        self.camera_height = self.camera_width / self.aspect_ratio
        self.cameraFwd = (look_at - look_from).normalize()
        self.cameraRight = self.cameraFwd.cross(vec3(0.0, 1.0, 0.0)).normalize()
        self.cameraUp = self.cameraRight.cross(self.cameraFwd)
        self.lens_radius = aperture / 2.0
        self.focal_distance = focal_distance
        self._sample_counter = 0

    def get_ray(self, n, scene=None, seed=0):
        """One jittered primary ray per pixel (row-major), as a Ray bundle, generated on the GPU.
        ``scene`` defaults to the scene this camera was attached to by ``Scene.add_Camera``."""
        from .ray import Ray
        scene = scene if scene is not None else getattr(self, "_scene", None)
        if scene is None:
            raise RuntimeError("Camera.get_ray needs the owning Scene (use Scene.add_Camera)")
        sample = self._sample_counter
        self._sample_counter += 1
        o, d = scene._backend().camera_rays(sample, seed)
        return Ray(vec3(o[:, 0], o[:, 1], o[:, 2]), vec3(d[:, 0], d[:, 1], d[:, 2]),
                   depth=0, n=n, reflections=0, transmissions=0, diffuse_reflections=0)
