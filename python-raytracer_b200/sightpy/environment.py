"""Backgrounds: a cube-map sky box or an equirectangular panorama, modelled (as upstream) as an
ordinary huge primitive with a self-emitting texture.

Reference: sightpy/backgrounds/skybox.py:9-94, sightpy/backgrounds/panorama.py:10-26.
"""
from .constants import SKYBOX_DISTANCE
from .imaging import DECODE_LINEAR, DECODE_PLAIN, TextureImage, open_rgb8
from .shading import Material
from .shapes import Cuboid_Collider, Primitive, Sphere_Collider
from .vec import vec3

__all__ = ["SkyBox", "Panorama", "SkyBox_Material"]


class SkyBox_Material(Material):
    """Environment lookup; adds ``light_intensity * lightmap`` for non-primary rays (skybox.py:35-94)."""

    def __init__(self, cubemap, light_intensity, blur):
        super().__init__()
        print("proccesing " + str(cubemap))
        raw = open_rgb8("sightpy/backgrounds/" + cubemap)
        self.texture = TextureImage(raw, DECODE_LINEAR)
        self.lightmap = None
        if light_intensity != 0.0:
            self.lightmap = TextureImage(open_rgb8("sightpy/backgrounds/lightmaps/" + cubemap), DECODE_PLAIN)
        self.blur_image = None
        if blur != 0.0:
            print("blurring " + str(cubemap))
            self.blur_image = TextureImage(raw, DECODE_LINEAR, cube_blur=blur, name=str(cubemap))
        self.blur = blur
        self.light_intensity = light_intensity
        self.repeat = 1.0


class SkyBox(Primitive):
    uv_cross_layout = True

    def __init__(self, cubemap, center=vec3(0.0, 0.0, 0.0), light_intensity=0.0, blur=0.0):
        super().__init__(center, SkyBox_Material(cubemap, light_intensity, blur), shadow=False)
        side = 2 * SKYBOX_DISTANCE
        self.light_intensity = light_intensity
        self.collider_list.append(Cuboid_Collider(
            assigned_primitive=self, center=center, width=side, height=side, length=side))


class Panorama(Primitive):
    def __init__(self, panorama, center=vec3(0.0, 0.0, 0.0), light_intensity=0.0, blur=0.0):
        super().__init__(center, SkyBox_Material(panorama, light_intensity, blur), shadow=False)
        self.light_intensity = light_intensity
        self.collider_list.append(Sphere_Collider(
            assigned_primitive=self, center=center, radius=SKYBOX_DISTANCE))
