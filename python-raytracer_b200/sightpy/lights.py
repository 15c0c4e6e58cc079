"""Light sources used by Glossy direct lighting (reference: sightpy/lights.py:7-52)."""
from .constants import SKYBOX_DISTANCE

__all__ = ["Light", "PointLight", "DirectionalLight"]


class Light:
    def __init__(self, pos, color):
        self.pos = pos
        self.color = color


class DirectionalLight(Light):
    """Light at infinity: direction ``Ldir`` (towards the light), irradiance ``color * N.L``;
    shadow distance SKYBOX_DISTANCE (lights.py:40-52)."""

    def __init__(self, Ldir, color):
        self.Ldir = Ldir
        self.color = color

    def get_L(self):
        return self.Ldir

    def get_distance(self, M):
        return SKYBOX_DISTANCE


class PointLight(Light):
    """Positional light, irradiance ``color * N.L / d^2 * 100``.  Upstream's ``get_L`` references
    undefined names (lights.py:30-31) so it never ran; here L = (pos - M)/|pos - M|."""
