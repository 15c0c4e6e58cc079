"""Shared scalar constants of the sightpy API (reference: sightpy/utils/constants.py:1-4).

On the device FARAWAY is represented by +inf / collider id -1 (1e39 overflows float32).
"""
UPWARDS = 1          # ray hits the outer face of a collider
UPDOWN = -1          # ray hits the inner face
FARAWAY = 1.0e39     # "no hit" distance of the numpy reference
SKYBOX_DISTANCE = 1.0e6
