"""Textures and material descriptions.

Reference: sightpy/textures/texture.py:9-39, sightpy/materials/{material,glossy,refractive,
thin_film_interference,diffuse,emissive}.py.  The classes carry the same constructor arguments
and attributes as the reference; the shading maths itself (``get_color`` upstream) is evaluated
by the CUDA backend (csrc/shade.cuh) — there is deliberately no CPU implementation here.
"""
from .imaging import DECODE_LINEAR, DECODE_PLAIN, TextureImage, open_rgb8
from .vec import vec3

__all__ = [
    "texture", "solid_color", "image",
    "Material", "Glossy", "Refractive", "ThinFilmInterference", "Diffuse", "Emissive",
]


# ---- textures ----------------------------------------------------------------------------------
class texture:
    """Base class of colour sources (texture.py:9-16)."""


class solid_color(texture):
    def __init__(self, color):
        self.color = color


class image(texture):
    """Nearest-neighbour, repeating image texture, linearised at load (texture.py:27-39)."""

    def __init__(self, img, repeat=1.0):
        print("proccesing " + str(img))
        self.img = TextureImage(open_rgb8("sightpy/textures/" + img), DECODE_LINEAR)
        self.repeat = repeat


def _as_texture(value, what):
    if isinstance(value, vec3):
        return solid_color(value)
    if isinstance(value, texture):
        return value
    raise TypeError(f"{what} must be a vec3/rgb colour or a texture, got {type(value).__name__}")


# ---- materials ---------------------------------------------------------------------------------
class Material:
    """Common part: optional tangent-space normal map (material.py:11-40)."""

    def __init__(self, normalmap=None):
        self.normalmap = None
        self.repeat = 1.0   # upstream forgets to set this when ``normalmap=`` is passed (AttributeError)
        if normalmap is not None:
            self.set_normalmap(normalmap)

    def set_normalmap(self, normalmap, repeat=1.0):
        self.normalmap = TextureImage(open_rgb8("sightpy/normalmaps/" + normalmap), DECODE_PLAIN)
        self.repeat = repeat

    def get_color(self, scene, ray, hit):  # pragma: no cover
        raise NotImplementedError(
            "materials are evaluated by the CUDA backend; python-level get_color overrides "
            "cannot run on the GPU (no CPU fallback by design)")


class Glossy(Material):
    """Lambert + Cook-Torrance(Phong lobe, Schlick Fresnel) + mirror reflection (glossy.py:11-110)."""

    def __init__(self, diff_color, roughness, spec_coeff, diff_coeff, n, **kwargs):
        super().__init__(**kwargs)
        self.diff_texture = _as_texture(diff_color, "diff_color")
        self.roughness = roughness
        self.spec_coeff = spec_coeff
        self.diff_coeff = diff_coeff
        self.n = n   # complex index of refraction per colour channel


class Refractive(Material):
    """Dielectric/absorbing medium with full complex Fresnel (refractive.py:10-123)."""

    def __init__(self, n, **kwargs):
        super().__init__(**kwargs)
        self.n = n


class ThinFilmInterference(Material):
    """Soap-bubble film: reflectance from a (cos theta, thickness) LUT (thin_film_interference.py:11-22)."""

    def __init__(self, thickness, noise=0.0, **kwargs):
        super().__init__(**kwargs)
        self.thickness = thickness
        self.thin_film_interference_reflectance = TextureImage(
            open_rgb8("sightpy/textures/thin_film_interference_n=1.4.png"), DECODE_PLAIN)
        self.thickness_noise = TextureImage(open_rgb8("sightpy/textures/noise.png"), DECODE_PLAIN)
        self.noise_factor = noise


class Diffuse(Material):
    """Lambertian Monte-Carlo estimator, 20 -> 1 -> 0 ray split (diffuse.py:12-124)."""

    def __init__(self, diff_color, diffuse_rays=20, ambient_weight=0.5, **kwargs):
        super().__init__(**kwargs)
        self.diff_texture = _as_texture(diff_color, "diff_color")
        self.diffuse_rays = diffuse_rays
        self.max_diffuse_reflections = 2
        self.ambient_weight = ambient_weight


class Emissive(Material):
    def __init__(self, color, **kwargs):
        self.texture_color = _as_texture(color, "color")
        super().__init__(**kwargs)
