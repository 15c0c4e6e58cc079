#!/usr/bin/env python
"""Two metallic spheres on a checkered floor under a cube-map sky (Whitted-style; cf. BASELINE.json config 1).

    PYTHONPATH=python-raytracer_b200 python examples/spheres.py
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "python-raytracer_b200"))
from sightpy import *  # noqa: F401,F403,E402

gold = Glossy(diff_color=rgb(1., .572, .184), n=vec3(0.15 + 3.58j, 0.4 + 2.37j, 1.54 + 1.91j), roughness=0.0,
              spec_coeff=0.2, diff_coeff=0.8)
blue = Glossy(diff_color=rgb(0.0, 0, 0.1), n=vec3(1.3 + 1.91j, 1.3 + 1.91j, 1.4 + 2.91j), roughness=0.2,
              spec_coeff=0.5, diff_coeff=0.3)
floor = Glossy(diff_color=image("checkered_floor.png", repeat=80.), n=vec3(1.2 + 0.3j, 1.2 + 0.3j, 1.1 + 0.3j),
               roughness=0.2, spec_coeff=0.3, diff_coeff=0.9)

Sc = Scene(ambient_color=rgb(0.05, 0.05, 0.05))
angle = -np.pi / 2 * 0.3
Sc.add_Camera(look_from=vec3(2.5 * np.sin(angle), 0.25, 2.5 * np.cos(angle) - 1.5), look_at=vec3(0., 0.25, -3.),
              screen_width=800, screen_height=600)
Sc.add_DirectionalLight(Ldir=vec3(0.52, 0.45, -0.5), color=rgb(0.15, 0.15, 0.15))
Sc.add(Sphere(material=gold, center=vec3(-.75, .1, -3.), radius=.6, max_ray_depth=3))
Sc.add(Sphere(material=blue, center=vec3(1.25, .1, -3.), radius=.6, max_ray_depth=3))
Sc.add(Plane(material=floor, center=vec3(0, -0.5, -3.0), width=120.0, height=120.0, u_axis=vec3(1.0, 0, 0),
             v_axis=vec3(0, 0, -1.0), max_ray_depth=3))
Sc.add_Background("stormydays.png")

img = Sc.render(samples_per_pixel=16)
img.save("spheres.png")
print(Sc.last_stats["rays_total"], "rays,", round(Sc.last_stats["device_ms"], 2), "ms on the device -> spheres.png")
