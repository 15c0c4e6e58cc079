"""Host-side image loading / colour-space helpers (one-off preprocessing; stays on the CPU).

Reference: sightpy/utils/image_functions.py:7-33, sightpy/utils/colour_functions.py:4-28,
sightpy/backgrounds/util/blur_background.py:17-132.

Every texture the renderer samples originates from an 8-bit image, and every float value the
reference derives from it is a pure function of that byte (``b/256`` or the sRGB->linear curve
of ``b/256``).  So textures are kept as uint8 RGB plus the *name* of a 256-entry decode table;
the GPU samples RGBA8 texels and decodes through the table (exact, 6x less HBM than float64x3).
"""
from pathlib import Path

import numpy as np
from PIL import Image, ImageFilter

__all__ = [
    "TextureImage", "decode_table", "resolve_asset", "open_rgb8",
    "load_image", "load_image_with_blur", "load_image_as_linear_sRGB",
    "sRGB_linear_to_sRGB", "sRGB_to_sRGB_linear", "blur_skybox_u8",
]

_PKG_DIR = Path(__file__).resolve().parent

DECODE_PLAIN = 0    # value = byte / 256                      (image_functions.py:7-9)
DECODE_LINEAR = 1   # value = sRGB_to_sRGB_linear(byte / 256) (image_functions.py:19-33)


def sRGB_linear_to_sRGB(rgb_linear):
    """sRGB OETF followed by per-pixel max-channel normalisation (colour_functions.py:4-18)."""
    rgb_linear = np.asarray(rgb_linear, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        encoded = np.where(rgb_linear <= 0.00304, 12.92 * rgb_linear,
                           1.055 * np.power(rgb_linear, 1.0 / 2.4) - 0.055)
    peak = np.amax(encoded, axis=0) + 0.00001
    return np.where(peak > 1.0, encoded / peak, encoded)


def sRGB_to_sRGB_linear(rgb):
    """Inverse sRGB transfer curve used when textures are loaded (colour_functions.py:21-28)."""
    rgb = np.asarray(rgb, dtype=np.float64)
    return np.where(rgb <= 0.03928, rgb / 12.92, np.power((rgb + 0.055) / 1.055, 2.4))


def decode_table(kind):
    """256-entry float64 table mapping a texel byte to the value the reference would hold."""
    b = np.arange(256, dtype=np.float64) / 256.0
    return sRGB_to_sRGB_linear(b) if kind == DECODE_LINEAR else b


_NEXT_TEXTURE_KEY = [1]


class TextureImage:
    """uint8 RGB image + decode-table kind. ``shape``/``as_float`` mimic the reference arrays.
    Immutable by convention: ``key`` names its bytes for the backend's device-resident texture cache.

    ``cube_blur`` > 0 marks a cross-layout cube map that is to be blurred (skybox.py:46-49).  The renderer hands the raw
    texels (``source_u8``) and the radius to the library, which blurs them on the GPU; ``u8`` — what host-side
    consumers such as the test oracle read — is then computed on demand with the reference's own Pillow pipeline
    (``blur_skybox_u8``), so the two can be compared byte for byte."""

    def __init__(self, u8, decode, cube_blur=0.0, name=""):
        self.key = _NEXT_TEXTURE_KEY[0]          # unique per object for the life of the process
        _NEXT_TEXTURE_KEY[0] += 1
        u8 = np.ascontiguousarray(u8, dtype=np.uint8)
        if u8.ndim != 3 or u8.shape[2] != 3:
            raise ValueError("TextureImage expects an H x W x 3 uint8 array")
        self.source_u8 = u8
        self.decode = int(decode)
        self.cube_blur = float(cube_blur)
        self.name = name
        self._u8 = None if self.cube_blur else u8

    @property
    def u8(self):
        if self._u8 is None:
            self._u8 = blur_skybox_u8(self.source_u8, self.cube_blur, self.name)
        return self._u8

    @property
    def shape(self):
        return self.source_u8.shape

    def as_float(self):
        return decode_table(self.decode)[self.u8]


def resolve_asset(relative):
    """Reference scripts rely on CWD-relative ``sightpy/<dir>/<file>`` paths (texture.py:29).
    Honour that first, then fall back to the copy that ships inside this package."""
    p = Path(relative)
    if p.exists():
        return p
    parts = p.parts
    if parts and parts[0] == "sightpy":
        q = _PKG_DIR.joinpath(*parts[1:])
        if q.exists():
            return q
    raise FileNotFoundError(f"asset not found: {relative}")


def open_rgb8(path, blur=0.0):
    img = Image.open(resolve_asset(path))
    if blur != 0.0:
        img = img.filter(ImageFilter.GaussianBlur(radius=blur))
    return np.asarray(img.convert("RGB"), dtype=np.uint8)


# -- API-compatible float loaders (kept for scripts that call them directly) -------------------
def load_image(path):
    return np.asarray(Image.open(resolve_asset(path))) / 256.0


def load_image_with_blur(path, blur=0.0):
    img = Image.open(resolve_asset(path)).filter(ImageFilter.GaussianBlur(radius=blur))
    return np.asarray(img) / 256.0


def load_image_as_linear_sRGB(path, blur=0.0):
    path = resolve_asset(path)
    print("proccesing " + str(path.name))
    img = Image.open(path)
    if blur != 0.0:
        img = img.filter(ImageFilter.GaussianBlur(radius=blur))
    return sRGB_to_sRGB_linear(np.asarray(img) / 256.0)


# -- cube-map blur -----------------------------------------------------------------------------
# For each face of the cross layout: the five tiles (left, centre, right, below, above) of the
# 3N x 3N canvas that is Gaussian-blurred so that the blur bleeds correctly across cube edges,
# as (source face, number of counter-clockwise quarter turns).  Restates the hand-unrolled
# canvases of blur_background.py:40-118.
_BLUR_NEIGHBOURS = {
    #  face      left           centre         right          below           above
    "back":   (("right", 0), ("back", 0),   ("left", 0),  ("bottom", 2), ("top", 2)),
    "top":    (("left", -1), ("top", 0),    ("right", 1), ("front", 0),  ("back", 2)),
    "bottom": (("left", 1),  ("bottom", 0), ("right", -1), ("back", 2),  ("front", 0)),
    "right":  (("front", 0), ("right", 0),  ("back", 0),  ("bottom", 1), ("top", -1)),
    "front":  (("left", 0),  ("front", 0),  ("right", 0), ("bottom", 0), ("top", 0)),
    "left":   (("back", 0),  ("left", 0),   ("front", 0), ("bottom", -1), ("top", 1)),
}
# (row block, column block) of each face inside the 3 x 4 cross image
_CROSS_SLOT = {"left": (1, 0), "front": (1, 1), "right": (1, 2), "back": (1, 3),
               "top": (0, 1), "bottom": (2, 1)}


def blur_skybox_u8(u8, blur, name=""):
    """Blur a cross-layout cube map face by face; returns the blurred cross as uint8 RGB.

    The reference works on ``byte/256`` floats and re-quantises each canvas with
    ``(255*x).astype(uint8)`` (blur_background.py:6-10), i.e. byte b becomes max(b-1, 0) before
    the Gaussian filter; the filtered bytes are then read back as ``byte/256`` and linearised.
    Returning the filtered bytes keeps that pipeline exact (decode with DECODE_LINEAR)."""
    n = int(u8.shape[0] / 3)
    requant = (255 * (u8.astype(np.float64) / 256.0)).astype(np.uint8)
    face = {k: requant[r * n:(r + 1) * n, c * n:(c + 1) * n] for k, (r, c) in _CROSS_SLOT.items()}
    out = np.zeros((3 * n, 4 * n, 3), dtype=np.uint8)
    canvas_slots = ((1, 0), (1, 1), (1, 2), (2, 1), (0, 1))
    for target, tiles in _BLUR_NEIGHBOURS.items():
        canvas = np.zeros((3 * n, 3 * n, 3), dtype=np.uint8)
        for (r, c), (src, turns) in zip(canvas_slots, tiles):
            canvas[r * n:(r + 1) * n, c * n:(c + 1) * n] = np.rot90(face[src], k=turns)
        # the reference filters three merged 'L' planes == an RGB image
        blurred = np.asarray(Image.fromarray(canvas, "RGB").filter(ImageFilter.GaussianBlur(radius=blur)))
        r, c = _CROSS_SLOT[target]
        out[r * n:(r + 1) * n, c * n:(c + 1) * n] = blurred[n:2 * n, n:2 * n]
    return out
