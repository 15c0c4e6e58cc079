"""Scene definitions used by tests, golden generation and the benchmark.

Each builder takes the API namespace ``ns`` (this repo's ``sightpy`` package, or the reference
package when generating golden vectors) and returns a populated Scene.  The geometry/material
parameters are those of the reference's example scripts (example1-4.py, example_cornellbox.py)
and of BASELINE.json's synthetic stress configuration.
"""
import numpy as np


def _checker_floor(ns, repeat, n, diff_coeff):
    return ns.Glossy(diff_color=ns.image("checkered_floor.png", repeat=repeat), n=n,
                     roughness=0.2, spec_coeff=0.3, diff_coeff=diff_coeff)


def _floor_plane(ns, material, size, depth):
    v = ns.vec3
    return ns.Plane(material=material, center=v(0, -0.5, -3.0), width=size, height=size,
                    u_axis=v(1.0, 0, 0), v_axis=v(0, 0, -1.0), max_ray_depth=depth)


def example1(ns, width=400, height=300):
    """Two metallic spheres on a checkered glossy floor under a cube-map sky."""
    v, rgb = ns.vec3, ns.rgb
    gold = ns.Glossy(diff_color=rgb(1.0, 0.572, 0.184), n=v(0.15 + 3.58j, 0.4 + 2.37j, 1.54 + 1.91j),
                     roughness=0.0, spec_coeff=0.2, diff_coeff=0.8)
    blue = ns.Glossy(diff_color=rgb(0.0, 0, 0.1), n=v(1.3 + 1.91j, 1.3 + 1.91j, 1.4 + 2.91j),
                     roughness=0.2, spec_coeff=0.5, diff_coeff=0.3)
    floor = _checker_floor(ns, 80.0, v(1.2 + 0.3j, 1.2 + 0.3j, 1.1 + 0.3j), 0.9)
    sc = ns.Scene(ambient_color=rgb(0.05, 0.05, 0.05))
    ang = -np.pi / 2 * 0.3
    sc.add_Camera(look_from=v(2.5 * np.sin(ang), 0.25, 2.5 * np.cos(ang) - 1.5), look_at=v(0.0, 0.25, -3.0),
                  screen_width=width, screen_height=height)
    sc.add_DirectionalLight(Ldir=v(0.52, 0.45, -0.5), color=rgb(0.15, 0.15, 0.15))
    sc.add(ns.Sphere(material=gold, center=v(-0.75, 0.1, -3.0), radius=0.6, max_ray_depth=3))
    sc.add(ns.Sphere(material=blue, center=v(1.25, 0.1, -3.0), radius=0.6, max_ray_depth=3))
    sc.add(_floor_plane(ns, floor, 120.0, 3))
    sc.add_Background("stormydays.png")
    return sc


def example2(ns, width=400, height=300, mc=False):
    """Three absorbing glass spheres (depth 3) over the checkered floor."""
    v, rgb = ns.vec3, ns.rgb
    glasses = [ns.Refractive(n=v(1.5 + 4e-8j, 1.5 + 4e-8j, 1.5 + 0.0j)),
               ns.Refractive(n=v(1.5 + 4e-8j, 1.5 + 0.0j, 1.5 + 4e-8j)),
               ns.Refractive(n=v(1.5 + 0.0j, 1.5 + 5e-8j, 1.5 + 5e-8j))]
    floor = _checker_floor(ns, 80.0, v(1.2 + 0.3j, 1.2 + 0.3j, 1.1 + 0.3j), 0.9)
    sc = ns.Scene(ambient_color=rgb(0.05, 0.05, 0.05))
    ang = np.pi / 2 * 0.3
    sc.add_Camera(look_from=v(2.5 * np.sin(ang), 0.25, 2.5 * np.cos(ang) - 1.5), look_at=v(0.0, 0.25, -1.5),
                  screen_width=width, screen_height=height)
    sc.add_DirectionalLight(Ldir=v(0.52, 0.45, -0.5), color=rgb(0.15, 0.15, 0.15))
    for x, g in zip((-1.2, 0.0, 1.2), glasses):
        sc.add(ns.Sphere(material=g, center=v(x, 0.0, -1.5), radius=0.5, shadow=False, max_ray_depth=3, mc=mc))
    sc.add(_floor_plane(ns, floor, 120.0, 3))
    sc.add_Background("miramar.jpeg")
    return sc


def example3(ns, width=400, height=300, normalmap=False):
    """Rotated glass cuboid (depth 5) on a coarse checkered floor; optional normal-mapped floor."""
    v, rgb = ns.vec3, ns.rgb
    floor = ns.Glossy(diff_color=ns.image("checkered_floor.png", repeat=2.0), roughness=0.2, spec_coeff=0.3,
                      diff_coeff=0.7, n=v(2.2, 2.2, 2.2))
    if normalmap:
        floor.set_normalmap("floor.jpg", repeat=4.0)
    glass = ns.Refractive(n=v(1.5 + 4e-8j, 1.5 + 0.0j, 1.5 + 4e-8j))
    sc = ns.Scene()
    sc.add_Camera(look_from=v(0.0, 0.25, 1.0), look_at=v(0.0, 0.25, -3.0), screen_width=width, screen_height=height)
    sc.add_DirectionalLight(Ldir=v(0.0, 0.5, 0.5), color=rgb(0.5, 0.5, 0.5))
    sc.add(_floor_plane(ns, floor, 6.0, 5))
    cb = ns.Cuboid(material=glass, center=v(0.00, 0.0001, -0.8), width=0.9, height=1.0, length=0.4,
                   shadow=False, max_ray_depth=5)
    cb.rotate(θ=30, u=v(0, 1, 0))
    sc.add(cb)
    sc.add_Background("stormydays.png")
    return sc


def example4(ns, width=400, height=300):
    """Thin-film soap bubble in front of a blurred, light-emitting sky box."""
    v, rgb = ns.vec3, ns.rgb
    sc = ns.Scene(ambient_color=rgb(0.01, 0.01, 0.01))
    ang = -np.pi * 0.5
    sc.add_Camera(screen_height=height, screen_width=width,
                  look_from=v(4.0 * np.sin(ang), 0.00, 4.0 * np.cos(ang)), look_at=v(0.0, 0.05, 0.0))
    bubble = ns.ThinFilmInterference(thickness=330, noise=60.0)
    sc.add(ns.Sphere(material=bubble, center=v(1.0, 0.0, 1.5), radius=1.7, shadow=False, max_ray_depth=5))
    sc.add_Background("lake.png", light_intensity=5.0, blur=10.0)
    return sc


def cornell(ns, width=100, height=100, mc=False):
    """Cornell box: emissive ceiling panel, diffuse walls, rotated tall box, glass sphere;
    light and sphere are importance sampled."""
    v, rgb = ns.vec3, ns.rgb
    sc = ns.Scene(ambient_color=rgb(0.00, 0.00, 0.00))
    sc.add_Camera(screen_width=width, screen_height=height, look_from=v(278, 278, 800), look_at=v(278, 278, 0),
                  focal_distance=1.0, field_of_view=40)
    green = ns.Diffuse(diff_color=rgb(0.12, 0.45, 0.15))
    red = ns.Diffuse(diff_color=rgb(0.65, 0.05, 0.05))
    white = ns.Diffuse(diff_color=rgb(0.73, 0.73, 0.73))
    lamp = ns.Emissive(color=rgb(15.0, 15.0, 15.0))
    glass = ns.Refractive(n=v(1.5 + 0.05e-8j, 1.5 + 0.02e-8j, 1.5 + 0.0j))
    sc.add(ns.Plane(material=lamp, center=v(213 + 130 / 2, 554, -227.0 - 105 / 2), width=130.0, height=105.0,
                    u_axis=v(1.0, 0.0, 0), v_axis=v(0.0, 0, 1.0)), importance_sampled=True)
    walls = [  # material, centre, u_axis, v_axis
        (white, v(555 / 2, 555 / 2, -555.0), v(0.0, 1.0, 0), v(1.0, 0, 0.0)),
        (green, v(-0.0, 555 / 2, -555 / 2), v(0.0, 1.0, 0), v(0.0, 0, -1.0)),
        (red, v(555.0, 555 / 2, -555 / 2), v(0.0, 1.0, 0), v(0.0, 0, -1.0)),
        (white, v(555 / 2, 555, -555 / 2), v(1.0, 0.0, 0), v(0.0, 0, -1.0)),
        (white, v(555 / 2, 0.0, -555 / 2), v(1.0, 0.0, 0), v(0.0, 0, -1.0)),
    ]
    for m, c, ua, va in walls:
        sc.add(ns.Plane(material=m, center=c, width=555.0, height=555.0, u_axis=ua, v_axis=va))
    box = ns.Cuboid(material=white, center=v(182.5, 165, -285 - 160 / 2), width=165, height=165 * 2, length=165,
                    shadow=False)
    box.rotate(θ=15, u=v(0, 1, 0))
    sc.add(box)
    sc.add(ns.Sphere(material=glass, center=v(370.5, 165 / 2, -65 - 185 / 2), radius=165 / 2, shadow=False,
                     max_ray_depth=3, mc=mc), importance_sampled=True)
    return sc


def triangles(ns, width=160, height=120):
    """Small scene exercising Triangle colliders (built directly, like SURVEY App. B prescribes for
    the reference), a panorama background and a point-free glossy set-up."""
    v, rgb = ns.vec3, ns.rgb
    sc = ns.Scene(ambient_color=rgb(0.1, 0.1, 0.1))
    sc.add_Camera(look_from=v(0.0, 1.0, 3.0), look_at=v(0.0, 0.5, 0.0), screen_width=width, screen_height=height,
                  field_of_view=60)
    sc.add_DirectionalLight(Ldir=v(0.3, 0.8, 0.5), color=rgb(0.6, 0.6, 0.6))
    mats = [ns.Glossy(diff_color=rgb(0.8, 0.2, 0.2), n=v(1.5 + 0.2j, 1.5 + 0.2j, 1.5 + 0.2j), roughness=0.3,
                      spec_coeff=0.4, diff_coeff=0.8),
            ns.Emissive(color=rgb(0.2, 0.9, 0.3)),
            ns.Refractive(n=v(1.4 + 1e-8j, 1.4 + 0j, 1.4 + 2e-8j))]
    tris = [(v(-1.5, 0.0, 0.0), v(0.0, 0.0, -0.5), v(-0.7, 1.6, -0.2)),
            (v(0.2, 0.1, 0.4), v(1.6, 0.0, -0.3), v(0.9, 1.4, 0.1)),
            (v(-0.6, 0.3, 1.0), v(0.7, 0.2, 1.2), v(0.0, 1.2, 0.9))]
    for m, (p1, p2, p3) in zip(mats, tris):
        prim = ns.Primitive(center=(p1 + p2 + p3) / 3, material=m, max_ray_depth=3, shadow=True)
        prim.collider_list += [ns.Triangle_Collider(assigned_surface=prim, p1=p1, p2=p2, p3=p3)]
        prim.bounded_sphere_radius = 1.0
        sc.add(prim)
    sc.add(ns.Plane(material=ns.Glossy(diff_color=rgb(0.5, 0.5, 0.6), n=v(1.3 + 0.1j, 1.3 + 0.1j, 1.3 + 0.1j),
                                       roughness=0.0, spec_coeff=0.3, diff_coeff=0.7),
                    center=v(0, -0.2, 0), width=12.0, height=12.0, u_axis=v(1.0, 0, 0), v_axis=v(0, 0, -1.0),
                    max_ray_depth=2))
    sc.add_Background("miramar.jpeg", spherical=True)
    return sc


def stress(ns, width=3840, height=2160, n_spheres=4096, n_triangles=1024, n_collections=2, seed=1234):
    """BASELINE.json config 5: random spheres + triangle collections over a checkered ground."""
    v, rgb = ns.vec3, ns.rgb
    rng = np.random.default_rng(seed)
    sc = ns.Scene(ambient_color=rgb(0.05, 0.05, 0.05))
    sc.add_Camera(look_from=v(0.0, 12.0, 30.0), look_at=v(0.0, 6.0, -60.0), screen_width=width,
                  screen_height=height, field_of_view=60)
    sc.add_DirectionalLight(Ldir=v(0.4, 0.8, 0.45), color=rgb(0.5, 0.5, 0.5))
    lo, hi = np.array([-60.0, 0.3, -140.0]), np.array([60.0, 25.0, -10.0])
    centres = rng.uniform(lo, hi, size=(n_spheres, 3))
    radii = rng.uniform(0.2, 1.2, size=n_spheres)
    albedo = rng.uniform(0.1, 0.9, size=(n_spheres, 3))
    for i in range(n_spheres):
        k = i % 10
        c = v(*map(float, centres[i]))
        if k < 6:
            m = ns.Diffuse(diff_color=rgb(*map(float, albedo[i])))
        elif k < 8:
            m = ns.Glossy(diff_color=rgb(*map(float, albedo[i])), n=v(1.5 + 1j, 1.5 + 1j, 1.5 + 1j), roughness=0.2,
                          spec_coeff=0.4, diff_coeff=0.7)
        elif k == 8:
            m = ns.Refractive(n=v(1.5 + 0j, 1.5 + 0j, 1.5 + 0j))
        else:
            m = ns.Emissive(color=rgb(4.0, 4.0, 4.0))
        sc.add(ns.Sphere(material=m, center=c, radius=float(radii[i]), max_ray_depth=3))
    for _ in range(n_collections):
        base = rng.uniform(lo, hi, size=(n_triangles, 3))
        offs = rng.uniform(-1.0, 1.0, size=(n_triangles, 3, 3))
        m = ns.Diffuse(diff_color=rgb(*map(float, rng.uniform(0.1, 0.9, size=3))))
        prim = ns.Primitive(center=v(0.0, 12.0, -75.0), material=m, max_ray_depth=3, shadow=True)
        for j in range(n_triangles):
            p = [v(*map(float, base[j] + offs[j, q])) for q in range(3)]
            prim.collider_list.append(ns.Triangle_Collider(assigned_surface=prim, p1=p[0], p2=p[1], p3=p[2]))
        prim.bounded_sphere_radius = 100.0
        sc.add(prim)
    ground = _checker_floor(ns, 40.0, v(1.2 + 0.3j, 1.2 + 0.3j, 1.1 + 0.3j), 0.9)
    sc.add(ns.Plane(material=ground, center=v(0, 0.0, -75.0), width=400.0, height=400.0, u_axis=v(1.0, 0, 0),
                    v_axis=v(0, 0, -1.0), max_ray_depth=2))
    return sc


def fuzz(ns, width=48, height=36, seed=0):
    """Seeded random Whitted scene: every collider type at random poses, every deterministic material with
    random parameters, random lights / background / depth limits.  Parameters are drawn first as plain
    floats so that the reference's classes and ours are handed identical numbers."""
    v, rgb = ns.vec3, ns.rgb
    rng = np.random.default_rng(1000 + seed)
    U = lambda lo, hi, *shape: rng.uniform(lo, hi, size=shape or None)   # noqa: E731
    f3 = lambda a: v(*map(float, a))                                    # noqa: E731

    def unit(a):
        a = np.asarray(a, dtype=np.float64)
        return a / np.linalg.norm(a)

    c3 = lambda a: v(*map(complex, a))                                  # noqa: E731

    def material(allow_emissive=True, uv=True):
        k = int(rng.integers(0, 5 if allow_emissive else 4))
        if k == 3 and not uv:
            k = 2       # thin films need a uv mapping, which triangles lack
        if k == 0:      # dielectric-ish glossy, optionally textured
            tex = uv and rng.random() < 0.4
            colour = ns.image("checkered_floor.png", repeat=float(U(1.0, 6.0))) if tex else rgb(*map(float, U(0.05, 0.95, 3)))
            return ns.Glossy(diff_color=colour, n=c3(U(1.1, 2.4, 3) + 1j * U(0.0, 0.6, 3)),
                             roughness=float(U(0.0, 0.6)), spec_coeff=float(U(0.1, 0.6)), diff_coeff=float(U(0.3, 0.9)))
        if k == 1:      # metal
            return ns.Glossy(diff_color=rgb(*map(float, U(0.05, 0.95, 3))), n=c3(U(0.1, 1.6, 3) + 1j * U(1.5, 3.8, 3)),
                             roughness=float(U(0.0, 0.3)), spec_coeff=float(U(0.2, 0.6)), diff_coeff=float(U(0.3, 0.9)))
        if k == 2:      # absorbing glass
            return ns.Refractive(n=c3(U(1.2, 1.8) + U(0.0, 0.05, 3) + 1j * U(0.0, 6e-8, 3)))
        if k == 3:
            return ns.ThinFilmInterference(thickness=float(U(150.0, 380.0)), noise=float(rng.choice([0.0, 20.0, 60.0])))
        return ns.Emissive(color=rgb(*map(float, U(0.2, 3.0, 3))))

    sc = ns.Scene(ambient_color=rgb(*map(float, U(0.0, 0.12, 3))))
    ang, rad = float(U(0, 2 * np.pi)), float(U(4.0, 6.5))
    sc.add_Camera(look_from=v(rad * np.sin(ang), float(U(0.4, 3.0)), rad * np.cos(ang)),
                  look_at=f3(U(-0.5, 0.5, 3) + np.array([0, 0.6, 0])), screen_width=width, screen_height=height,
                  field_of_view=float(U(45.0, 85.0)))
    for _ in range(int(rng.integers(1, 3))):
        sc.add_DirectionalLight(Ldir=f3(unit(U(-1, 1, 3) + np.array([0, 1.2, 0]))), color=rgb(*map(float, U(0.1, 0.6, 3))))
    # ground: slightly tilted bounded plane
    up = unit(np.array([0, 1.0, 0]) + U(-0.08, 0.08, 3))
    ua = unit(np.cross(up, [0.0, 0.0, 1.0])); va = np.cross(ua, up)
    sc.add(ns.Plane(material=material(False), center=v(0.0, -0.6, 0.0), width=float(U(8, 30)), height=float(U(8, 30)),
                    u_axis=f3(ua), v_axis=f3(va), max_ray_depth=int(rng.integers(1, 5)), shadow=bool(rng.random() < 0.8)))
    for _ in range(int(rng.integers(2, 5))):
        sc.add(ns.Sphere(material=material(), center=f3(U(-2.2, 2.2, 3) * np.array([1, 0.5, 1]) + np.array([0, 0.5, 0])),
                         radius=float(U(0.25, 0.9)), max_ray_depth=int(rng.integers(1, 6)), shadow=bool(rng.random() < 0.7)))
    for _ in range(int(rng.integers(1, 3))):
        cb = ns.Cuboid(material=material(), center=f3(U(-2.0, 2.0, 3) * np.array([1, 0.4, 1]) + np.array([0, 0.4, 0])),
                       width=float(U(0.3, 1.2)), height=float(U(0.3, 1.4)), length=float(U(0.3, 1.2)),
                       shadow=bool(rng.random() < 0.7), max_ray_depth=int(rng.integers(1, 6)))
        for _ in range(int(rng.integers(1, 3))):
            cb.rotate(θ=float(U(0, 360)), u=f3(unit(U(-1, 1, 3))))
        sc.add(cb)
    for _ in range(int(rng.integers(0, 4))):
        c = U(-2.0, 2.0, 3) * np.array([1, 0.4, 1]) + np.array([0, 0.8, 0])
        p = [f3(c + U(-1.0, 1.0, 3)) for _ in range(3)]
        prim = ns.Primitive(center=f3(c), material=material(uv=False), max_ray_depth=int(rng.integers(1, 4)),
                            shadow=bool(rng.random() < 0.7))
        prim.collider_list += [ns.Triangle_Collider(assigned_surface=prim, p1=p[0], p2=p[1], p3=p[2])]
        prim.bounded_sphere_radius = 2.0
        sc.add(prim)
    if rng.random() < 0.6:      # free-standing bounded plane at a random pose, then rotated as a primitive
        nrm = unit(U(-1, 1, 3)); ua = unit(np.cross(nrm, unit(U(-1, 1, 3)))); va = np.cross(nrm, ua)
        pl = ns.Plane(material=material(), center=f3(U(-2.5, 2.5, 3) * np.array([1, 0.3, 1]) + np.array([0, 1.0, 0])),
                      width=float(U(0.8, 3.0)), height=float(U(0.8, 3.0)), u_axis=f3(ua), v_axis=f3(va),
                      max_ray_depth=int(rng.integers(1, 5)), shadow=bool(rng.random() < 0.7))
        if rng.random() < 0.5:
            pl.rotate(θ=float(U(0, 360)), u=f3(unit(U(-1, 1, 3))))
        sc.add(pl)
    bg = int(rng.integers(0, 4))
    if bg == 0:
        sc.add_Background("stormydays.png")
    elif bg == 1:
        sc.add_Background("miramar.jpeg", spherical=True)
    elif bg == 2:
        sc.add_Background("lake.png", light_intensity=float(U(0.5, 3.0)), blur=0.0)
    return sc


def fuzzmc(ns, width=32, height=24, seed=0):
    """Seeded random Monte-Carlo scene: a closed-ish room of Diffuse surfaces (solid and textured, several
    fan sizes and ambient weights), emitters of every collider type, glass / glossy / thin-film bystanders,
    0-3 importance-sampled primitives."""
    v, rgb = ns.vec3, ns.rgb
    rng = np.random.default_rng(5000 + seed)
    U = lambda lo, hi, *shape: rng.uniform(lo, hi, size=shape or None)   # noqa: E731
    f3 = lambda a: v(*map(float, a))                                    # noqa: E731
    c3 = lambda a: v(*map(complex, a))                                  # noqa: E731

    def unit(a):
        a = np.asarray(a, dtype=np.float64)
        return a / np.linalg.norm(a)

    fans = [20] + [int(x) for x in rng.choice([4, 7, 12], size=int(rng.integers(0, 3)), replace=False)]

    def diffuse(uv=True):
        tex = uv and rng.random() < 0.3
        colour = ns.image("checkered_floor.png", repeat=float(U(1.0, 5.0))) if tex else rgb(*map(float, U(0.1, 0.9, 3)))
        return ns.Diffuse(diff_color=colour, diffuse_rays=int(rng.choice(fans)), ambient_weight=float(rng.choice([0.5, 0.3, 0.8])))

    def bystander():
        k = int(rng.integers(0, 4))
        if k == 0:
            return ns.Refractive(n=c3(U(1.3, 1.7) + np.zeros(3) + 1j * U(0.0, 3e-8, 3)))
        if k == 1:
            return ns.Glossy(diff_color=rgb(*map(float, U(0.1, 0.9, 3))), n=c3(U(1.1, 2.0, 3) + 1j * U(0.0, 2.0, 3)),
                             roughness=float(U(0.0, 0.4)), spec_coeff=float(U(0.1, 0.5)), diff_coeff=float(U(0.3, 0.9)))
        if k == 2:
            return ns.ThinFilmInterference(thickness=float(U(150.0, 380.0)), noise=float(rng.choice([0.0, 40.0])))
        return diffuse()

    sc = ns.Scene(ambient_color=rgb(*map(float, U(0.0, 0.05, 3))))
    sc.add_Camera(look_from=v(float(U(-0.5, 0.5)), float(U(1.5, 2.5)), 7.5), look_at=v(0.0, 2.0, 0.0), screen_width=width,
                  screen_height=height, field_of_view=float(U(40.0, 60.0)))
    if rng.random() < 0.5:
        sc.add_DirectionalLight(Ldir=f3(unit(U(-1, 1, 3) + np.array([0, 1.5, 0]))), color=rgb(*map(float, U(0.1, 0.4, 3))))
    n_importance = int(rng.integers(0, 4))
    # room: floor, back wall, two side walls, ceiling (4 x 4 x 4, open towards the camera)
    walls = [((0, 0, 0), (1, 0, 0), (0, 0, -1)), ((0, 2, -2), (1, 0, 0), (0, 1, 0)), ((-2, 2, 0), (0, 0, 1), (0, 1, 0)),
             ((2, 2, 0), (0, 0, -1), (0, 1, 0)), ((0, 4, 0), (1, 0, 0), (0, 0, 1))]
    for c, ua, va in walls:
        sc.add(ns.Plane(material=diffuse(), center=f3(c), width=4.0, height=4.0, u_axis=f3(ua), v_axis=f3(va),
                        max_ray_depth=int(rng.integers(2, 5))))
    # emitters
    kind = int(rng.integers(0, 3))
    lamp = ns.Emissive(color=rgb(*map(float, U(4.0, 16.0, 3))))
    if kind == 0:
        em = ns.Plane(material=lamp, center=v(float(U(-0.8, 0.8)), 3.98, float(U(-0.8, 0.8))), width=float(U(0.6, 1.6)),
                      height=float(U(0.6, 1.6)), u_axis=v(1.0, 0, 0), v_axis=v(0, 0, 1.0))
    elif kind == 1:
        em = ns.Sphere(material=lamp, center=f3(U(-1.2, 1.2, 3) + np.array([0, 2.8, 0])), radius=float(U(0.15, 0.45)))
    else:
        em = ns.Cuboid(material=lamp, center=f3(U(-1.0, 1.0, 3) + np.array([0, 3.0, 0])), width=float(U(0.3, 0.8)),
                       height=float(U(0.2, 0.5)), length=float(U(0.3, 0.8)))
        em.rotate(θ=float(U(0, 90)), u=f3(unit(U(-1, 1, 3))))
    sc.add(em, importance_sampled=n_importance >= 1)
    for j in range(int(rng.integers(2, 5))):
        c = U(-1.4, 1.4, 3) + np.array([0, 1.2, 0])
        m = bystander()
        if rng.random() < 0.6:
            ob = ns.Sphere(material=m, center=f3(c), radius=float(U(0.3, 0.7)), max_ray_depth=int(rng.integers(2, 5)),
                           shadow=bool(rng.random() < 0.5), mc=bool(rng.random() < 0.5))
        else:
            ob = ns.Cuboid(material=m, center=f3(c), width=float(U(0.4, 1.0)), height=float(U(0.4, 1.6)),
                           length=float(U(0.4, 1.0)), max_ray_depth=int(rng.integers(2, 5)), shadow=bool(rng.random() < 0.5))
            ob.rotate(θ=float(U(0, 360)), u=f3(unit(U(-0.3, 0.3, 3) + np.array([0, 1.0, 0]))))
        sc.add(ob, importance_sampled=(j + 2) <= n_importance)
    if rng.random() < 0.5:
        c = U(-1.0, 1.0, 3) + np.array([0, 2.0, 0])
        prim = ns.Primitive(center=f3(c), material=diffuse(uv=False), max_ray_depth=3, shadow=True)
        prim.collider_list += [ns.Triangle_Collider(assigned_surface=prim, p1=f3(c + U(-1, 1, 3)), p2=f3(c + U(-1, 1, 3)),
                                                    p3=f3(c + U(-1, 1, 3)))]
        prim.bounded_sphere_radius = 2.0
        sc.add(prim)
    if rng.random() < 0.5:
        sc.add_Background("lake.png", light_intensity=float(U(0.5, 2.0)), blur=0.0)
    return sc


BUILDERS = {"example1": example1, "example2": example2, "example3": example3, "example4": example4,
            "cornell": cornell, "triangles": triangles, "stress": stress, "fuzz": fuzz, "fuzzmc": fuzzmc}
