"""sightpy — the Python-Raytracer API on a B200-native CUDA backend.

Star-exports the same public names as the reference package (sightpy/__init__.py:1-12), including
``np``, ``Image`` and friends that its example scripts pick up through ``from sightpy import *``.
"""
import numbers  # noqa: F401
from multiprocessing import Pool, cpu_count  # noqa: F401  (re-exported by the reference too)

import numpy as np  # noqa: F401
from PIL import Image  # noqa: F401

from .constants import *  # noqa: F401,F403
from .vec import *  # noqa: F401,F403
from .imaging import (load_image, load_image_with_blur, load_image_as_linear_sRGB,  # noqa: F401
                      sRGB_linear_to_sRGB, sRGB_to_sRGB_linear)
from .ray import *  # noqa: F401,F403
from .camera import *  # noqa: F401,F403
from .shapes import *  # noqa: F401,F403
from .lights import *  # noqa: F401,F403
from .shading import *  # noqa: F401,F403
from .environment import *  # noqa: F401,F403
from .scene import *  # noqa: F401,F403
from .animation import *  # noqa: F401,F403
