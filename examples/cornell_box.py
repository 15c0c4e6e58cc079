#!/usr/bin/env python
"""Cornell box with the sightpy API on the B200 backend (scene of BASELINE.json config 4).

    PYTHONPATH=python-raytracer_b200 python examples/cornell_box.py [width height spp]
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "python-raytracer_b200"))
from sightpy import *  # noqa: F401,F403,E402  (the reference's scripts start the same way)

width, height, spp = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (640, 360, 64)

Sc = Scene(ambient_color=rgb(0.00, 0.00, 0.00))
Sc.add_Camera(screen_width=width, screen_height=height, look_from=vec3(278, 278, 800), look_at=vec3(278, 278, 0),
              focal_distance=1.0, field_of_view=40)

green, red, white = (Diffuse(diff_color=c) for c in (rgb(.12, .45, .15), rgb(.65, .05, .05), rgb(.73, .73, .73)))
lamp = Emissive(color=rgb(15.0, 15.0, 15.0))
glass = Refractive(n=vec3(1.5 + 0.05e-8j, 1.5 + 0.02e-8j, 1.5 + 0.j))

Sc.add(Plane(material=lamp, center=vec3(213 + 130 / 2, 554, -227.0 - 105 / 2), width=130.0, height=105.0,
             u_axis=vec3(1.0, 0.0, 0), v_axis=vec3(0.0, 0, 1.0)), importance_sampled=True)
for material, center, u_axis, v_axis in (
        (white, vec3(555 / 2, 555 / 2, -555.0), vec3(0.0, 1.0, 0), vec3(1.0, 0, 0.0)),      # back
        (green, vec3(-0.0, 555 / 2, -555 / 2), vec3(0.0, 1.0, 0), vec3(0.0, 0, -1.0)),       # left
        (red, vec3(555.0, 555 / 2, -555 / 2), vec3(0.0, 1.0, 0), vec3(0.0, 0, -1.0)),        # right
        (white, vec3(555 / 2, 555, -555 / 2), vec3(1.0, 0.0, 0), vec3(0.0, 0, -1.0)),        # ceiling
        (white, vec3(555 / 2, 0.0, -555 / 2), vec3(1.0, 0.0, 0), vec3(0.0, 0, -1.0))):       # floor
    Sc.add(Plane(material=material, center=center, width=555.0, height=555.0, u_axis=u_axis, v_axis=v_axis))

box = Cuboid(material=white, center=vec3(182.5, 165, -285 - 160 / 2), width=165, height=165 * 2, length=165, shadow=False)
box.rotate(θ=15, u=vec3(0, 1, 0))
Sc.add(box)
Sc.add(Sphere(material=glass, center=vec3(370.5, 165 / 2, -65 - 185 / 2), radius=165 / 2, shadow=False, max_ray_depth=3),
       importance_sampled=True)

img = Sc.render(samples_per_pixel=spp)
img.save("cornell_box.png")
s = Sc.last_stats
print(f"{s['rays_total'] / 1e6:.1f} M rays in {s['device_ms']:.1f} ms on the device "
      f"({s['rays_total'] / s['device_ms'] / 1e6:.2f} Grays/s) -> cornell_box.png")
