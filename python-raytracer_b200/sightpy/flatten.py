"""Scene objects -> plain-old-data arrays (the ``FlatScene``).

This is the marshalling step of the drop-in boundary: everything ``Scene.render`` needs is
reduced to a handful of C-layout numpy records (``include/sightpy_b200.h`` declares the same
structs) plus a list of 8-bit textures.  The CUDA library and the float64 test oracle both
consume exactly this description, so the flattening itself is covered by the parity tests.

Objects are read by attribute name only ("duck typing"), which lets the golden-vector generator
flatten scenes built from the *reference's* classes as well (tests/golden/make_golden.py).
Reference for the scene bookkeeping: sightpy/scene.py:29-69.
"""
from dataclasses import dataclass, field

import numpy as np

from .imaging import DECODE_LINEAR, DECODE_PLAIN, TextureImage, decode_table

# ---- enums shared with include/sightpy_b200.h ---------------------------------------------------
COLLIDER_SPHERE, COLLIDER_PLANE, COLLIDER_CUBOID, COLLIDER_TRIANGLE = 0, 1, 2, 3
MAT_GLOSSY, MAT_REFRACTIVE, MAT_THINFILM, MAT_DIFFUSE, MAT_EMISSIVE, MAT_SKYBOX = 0, 1, 2, 3, 4, 5
LIGHT_DIRECTIONAL, LIGHT_POINT = 0, 1
COLLIDER_PAYLOAD = 40

CAMERA_DT = np.dtype([
    ("look_from", "f8", 3), ("right", "f8", 3), ("up", "f8", 3), ("fwd", "f8", 3),
    ("cam_w", "f8"), ("cam_h", "f8"), ("lens_radius", "f8"), ("focal_distance", "f8"),
    ("width", "i4"), ("height", "i4"),
], align=True)

MATERIAL_DT = np.dtype([
    ("kind", "i4"), ("medium", "i4"), ("normalmap_tex", "i4"), ("color_tex", "i4"),
    ("aux_tex0", "i4"), ("aux_tex1", "i4"), ("diffuse_rays", "i4"), ("max_diffuse_reflections", "i4"),
    ("index_h", "i4"), ("index_w", "i4"),
    ("normalmap_repeat", "f8"), ("color_repeat", "f8"), ("color", "f8", 3),
    ("n_re", "f8", 3), ("n_im", "f8", 3),
    ("roughness", "f8"), ("spec_coeff", "f8"), ("diff_coeff", "f8"),
    ("thickness", "f8"), ("noise_factor", "f8"), ("ambient_weight", "f8"), ("light_intensity", "f8"),
], align=True)

PRIMITIVE_DT = np.dtype([
    ("material", "i4"), ("max_ray_depth", "i4"), ("shadow", "i4"), ("mc", "i4"),
    ("uv_cross_layout", "i4"), ("_pad", "i4"),
    ("center", "f8", 3), ("bounded_sphere_radius", "f8"),
], align=True)

COLLIDER_DT = np.dtype([
    ("type", "i4"), ("primitive", "i4"), ("p", "f8", COLLIDER_PAYLOAD),
], align=True)

LIGHT_DT = np.dtype([
    ("kind", "i4"), ("_pad", "i4"), ("vec", "f8", 3), ("color", "f8", 3),
], align=True)

# payload slots of COLLIDER_DT["p"] per collider type (mirrors the union in the header)
SPHERE_SLOTS = {"center": (0, 3), "radius": (3, 4)}
PLANE_SLOTS = {"center": (0, 3), "u_axis": (3, 6), "v_axis": (6, 9), "normal": (9, 12), "w": (12, 13),
               "h": (13, 14), "uv_shift": (14, 16), "inv_basis": (16, 25)}
CUBOID_SLOTS = {"center": (0, 3), "ax_w": (3, 6), "ax_h": (6, 9), "ax_l": (9, 12), "lb_local": (12, 15),
                "rt_local": (15, 18), "size": (18, 21), "basis": (21, 30), "inv_basis": (30, 39)}
TRIANGLE_SLOTS = {"p1": (0, 3), "p2": (3, 6), "p3": (6, 9), "normal": (9, 12), "centroid": (12, 15),
                  "n31": (15, 18), "n12": (18, 21), "n23": (21, 24)}
SLOTS = {COLLIDER_SPHERE: SPHERE_SLOTS, COLLIDER_PLANE: PLANE_SLOTS,
         COLLIDER_CUBOID: CUBOID_SLOTS, COLLIDER_TRIANGLE: TRIANGLE_SLOTS}


@dataclass
class FlatScene:
    ambient: np.ndarray                      # (3,) f8
    camera: np.ndarray                       # CAMERA_DT scalar record (shape ())
    media: np.ndarray                        # (n_media, 3) c16; row 0 = scene.n
    materials: np.ndarray                    # MATERIAL_DT[n]
    primitives: np.ndarray                   # PRIMITIVE_DT[n]
    colliders: np.ndarray                    # COLLIDER_DT[n]  (ORDER == scene.collider_list)
    lights: np.ndarray                       # LIGHT_DT[n]
    importance: np.ndarray                   # i4[n] primitive ids
    shadow_colliders: np.ndarray             # i4[n] collider ids
    textures: list = field(default_factory=list)   # [TextureImage]

    def field_of(self, ci, name):
        lo, hi = SLOTS[int(self.colliders["type"][ci])][name]
        v = self.colliders["p"][ci, lo:hi]
        return v[0] if hi - lo == 1 else v

    def max_depth_bound(self):
        """Upper bound on ``ray.depth`` of any ray that can still spawn children: specular bounces
        stop at the largest ``max_ray_depth``; every Diffuse hit beyond that needs a fresh diffuse
        budget, of which a path has at most max_diffuse_reflections."""
        d = int(self.primitives["max_ray_depth"].max()) if len(self.primitives) else 0
        dif = self.materials["kind"] == MAT_DIFFUSE
        # a first Diffuse hit always fans out, whatever max_diffuse_reflections says (diffuse.py:34)
        extra = max(int(self.materials["max_diffuse_reflections"][dif].max()), 1) if dif.any() else 0
        return d + extra + 1


# ---- helpers ----------------------------------------------------------------------------------
def _v3(v):
    return np.array([v.x, v.y, v.z], dtype=np.float64)

def _c3(v):
    return np.array([v.x, v.y, v.z], dtype=np.complex128)

def _kind(obj):
    return type(obj).__name__

def texture_from_float(arr, decode):
    """Recover the uint8 texture behind one of the reference's float images (exact inverse of the
    decode table).  Only used when flattening reference-built scenes for golden vectors."""
    arr = np.asarray(arr, dtype=np.float64)[..., :3]
    table = decode_table(decode)
    idx = np.clip(np.searchsorted(table, arr), 0, 255)
    lower = np.clip(idx - 1, 0, 255)
    idx = np.where(np.abs(table[lower] - arr) < np.abs(table[idx] - arr), lower, idx)
    if not np.array_equal(table[idx], arr):
        raise ValueError("float image is not an exact decode of 8-bit data")
    return TextureImage(idx.astype(np.uint8), decode)


class _Flattener:
    def __init__(self):
        self.textures, self._tex_ids = [], {}
        self.media, self._medium_ids = [], {}
        self.materials, self._mat_ids = [], {}
        self.primitives, self._prim_ids = [], {}

    # textures -----------------------------------------------------------------------------
    def tex(self, img, decode):
        if img is None:
            return -1
        key = id(img)
        if key not in self._tex_ids:
            t = img if isinstance(img, TextureImage) else texture_from_float(img, decode)
            if t.decode != decode:
                raise ValueError("texture decode kind mismatch")
            self._tex_ids[key] = len(self.textures)
            self.textures.append(t)
            self._keep = getattr(self, "_keep", []) + [img]   # keep ids alive
        return self._tex_ids[key]

    def color_source(self, rec, tex_obj):
        """solid_color / image -> (color_tex, color_repeat, color)."""
        if _kind(tex_obj) == "solid_color":
            rec["color_tex"] = -1
            rec["color"] = _v3(tex_obj.color)
        elif _kind(tex_obj) == "image":
            rec["color_tex"] = self.tex(tex_obj.img, DECODE_LINEAR)
            rec["color_repeat"] = tex_obj.repeat
        else:
            raise TypeError(f"unsupported texture type {_kind(tex_obj)} (the GPU backend cannot run "
                            "python texture subclasses)")

    # media ----------------------------------------------------------------------------------
    def medium(self, n):
        key = tuple(_c3(n))
        if key not in self._medium_ids:
            self._medium_ids[key] = len(self.media)
            self.media.append(np.array(key, dtype=np.complex128))
        return self._medium_ids[key]

    # materials ------------------------------------------------------------------------------
    def material(self, m):
        if id(m) in self._mat_ids:
            return self._mat_ids[id(m)]
        rec = np.zeros((), dtype=MATERIAL_DT)
        for f in ("medium", "normalmap_tex", "color_tex", "aux_tex0", "aux_tex1"):
            rec[f] = -1
        rec["color_repeat"] = 1.0
        k = _kind(m)
        nm = getattr(m, "normalmap", None)
        if nm is not None:
            rec["normalmap_tex"] = self.tex(nm, DECODE_PLAIN)
            rec["normalmap_repeat"] = m.repeat
        if k == "Glossy":
            rec["kind"] = MAT_GLOSSY
            self.color_source(rec, m.diff_texture)
            n = _c3(m.n)
            rec["n_re"], rec["n_im"] = n.real, n.imag
            rec["roughness"], rec["spec_coeff"], rec["diff_coeff"] = m.roughness, m.spec_coeff, m.diff_coeff
        elif k == "Refractive":
            rec["kind"] = MAT_REFRACTIVE
            n = _c3(m.n)
            rec["n_re"], rec["n_im"] = n.real, n.imag
            rec["medium"] = self.medium(m.n)
        elif k == "ThinFilmInterference":
            rec["kind"] = MAT_THINFILM
            rec["thickness"], rec["noise_factor"] = m.thickness, m.noise_factor
            rec["aux_tex0"] = self.tex(m.thin_film_interference_reflectance, DECODE_PLAIN)
            noise = m.thickness_noise
            if not isinstance(noise, TextureImage):   # reference keeps channel 0 only
                noise = self._noise_cache(noise)
            rec["aux_tex1"] = self.tex(noise, DECODE_PLAIN)
        elif k == "Diffuse":
            rec["kind"] = MAT_DIFFUSE
            self.color_source(rec, m.diff_texture)
            rec["diffuse_rays"] = m.diffuse_rays
            rec["max_diffuse_reflections"] = m.max_diffuse_reflections
            rec["ambient_weight"] = m.ambient_weight
        elif k == "Emissive":
            rec["kind"] = MAT_EMISSIVE
            self.color_source(rec, m.texture_color)
        elif k == "SkyBox_Material":
            rec["kind"] = MAT_SKYBOX
            env = self.tex(m.texture, DECODE_LINEAR)
            rec["index_h"], rec["index_w"] = self.textures[env].shape[:2]
            blur_img = getattr(m, "blur_image", None) if m.blur != 0.0 else None
            rec["color_tex"] = self.tex(blur_img, DECODE_LINEAR) if blur_img is not None else env
            rec["color_repeat"] = m.repeat
            rec["light_intensity"] = m.light_intensity
            if m.light_intensity != 0.0:
                rec["aux_tex0"] = self.tex(m.lightmap, DECODE_PLAIN)
        else:
            raise TypeError(
                f"material {k} is not supported by the CUDA backend (Glossy, Refractive, "
                "ThinFilmInterference, Diffuse, Emissive and backgrounds are); python Material "
                "subclasses cannot run on the GPU and there is no CPU fallback")
        self._mat_ids[id(m)] = len(self.materials)
        self.materials.append(rec)
        return self._mat_ids[id(m)]

    def _noise_cache(self, arr2d):
        cache = self.__dict__.setdefault("_noise", {})
        if id(arr2d) not in cache:
            rgb = np.repeat(np.asarray(arr2d)[..., None], 3, axis=2)
            cache[id(arr2d)] = (arr2d, texture_from_float(rgb, DECODE_PLAIN))
        return cache[id(arr2d)][1]

    # primitives -----------------------------------------------------------------------------
    def primitive(self, p):
        if id(p) in self._prim_ids:
            return self._prim_ids[id(p)]
        rec = np.zeros((), dtype=PRIMITIVE_DT)
        rec["material"] = self.material(p.material)
        rec["max_ray_depth"] = p.max_ray_depth
        rec["shadow"] = bool(p.shadow)
        rec["mc"] = bool(getattr(p, "mc", False))
        rec["uv_cross_layout"] = _kind(p) in ("Cuboid", "SkyBox") or bool(getattr(p, "uv_cross_layout", False))
        rec["center"] = _v3(p.center)
        rec["bounded_sphere_radius"] = getattr(p, "bounded_sphere_radius", 0.0)
        self._prim_ids[id(p)] = len(self.primitives)
        self.primitives.append(rec)
        return self._prim_ids[id(p)]

    # colliders ------------------------------------------------------------------------------
    def collider(self, c):
        rec = np.zeros((), dtype=COLLIDER_DT)
        rec["primitive"] = self.primitive(c.assigned_primitive)
        p = rec["p"]
        k = _kind(c)

        def put(slots, name, value):
            lo, hi = slots[name]
            p[lo:hi] = np.asarray(value, dtype=np.float64).reshape(-1)

        if k == "Sphere_Collider":
            rec["type"] = COLLIDER_SPHERE
            put(SPHERE_SLOTS, "center", _v3(c.center)); put(SPHERE_SLOTS, "radius", c.radius)
        elif k == "Plane_Collider":
            rec["type"] = COLLIDER_PLANE
            for name in ("center", "u_axis", "v_axis", "normal"):
                put(PLANE_SLOTS, name, _v3(getattr(c, name)))
            put(PLANE_SLOTS, "w", c.w); put(PLANE_SLOTS, "h", c.h)
            put(PLANE_SLOTS, "uv_shift", c.uv_shift)
            put(PLANE_SLOTS, "inv_basis", c.inverse_basis_matrix)
        elif k == "Cuboid_Collider":
            rec["type"] = COLLIDER_CUBOID
            for name in ("center", "ax_w", "ax_h", "ax_l"):
                put(CUBOID_SLOTS, name, _v3(getattr(c, name)))
            put(CUBOID_SLOTS, "lb_local", _v3(c.lb_local_basis))
            put(CUBOID_SLOTS, "rt_local", _v3(c.rt_local_basis))
            put(CUBOID_SLOTS, "size", [c.width, c.height, c.length])
            put(CUBOID_SLOTS, "basis", c.basis_matrix)
            put(CUBOID_SLOTS, "inv_basis", c.inverse_basis_matrix)
        elif k == "Triangle_Collider":
            rec["type"] = COLLIDER_TRIANGLE
            for name in ("p1", "p2", "p3", "normal", "centroid", "n31", "n12", "n23"):
                put(TRIANGLE_SLOTS, name, _v3(getattr(c, name)))
        else:
            raise TypeError(
                f"collider {k} is not supported by the CUDA backend (sphere, plane, cuboid and "
                "triangle are); python Collider subclasses cannot run on the GPU")
        return rec


def flatten_scene(scene):
    """Flatten a Scene (ours or the reference's) into a FlatScene."""
    fl = _Flattener()
    fl.medium(scene.n)                                   # medium 0 = the scene's ambient medium
    colliders = [fl.collider(c) for c in scene.collider_list]
    coll_index = {id(c): i for i, c in enumerate(scene.collider_list)}
    for p in scene.importance_sampled_list:
        fl.primitive(p)
    flat = FlatScene(
        ambient=_v3(scene.ambient_color),
        camera=flatten_camera(scene.camera) if getattr(scene, "camera", None) is not None else np.zeros((), CAMERA_DT),
        media=np.array(fl.media, dtype=np.complex128).reshape(-1, 3),
        materials=np.array(fl.materials, dtype=MATERIAL_DT).reshape(-1),
        primitives=np.array(fl.primitives, dtype=PRIMITIVE_DT).reshape(-1),
        colliders=np.array(colliders, dtype=COLLIDER_DT).reshape(-1),
        lights=np.array([flatten_light(l) for l in scene.Light_list], dtype=LIGHT_DT).reshape(-1),
        importance=np.array([fl.primitive(p) for p in scene.importance_sampled_list], dtype=np.int32),
        shadow_colliders=np.array([coll_index[id(c)] for c in scene.shadowed_collider_list], dtype=np.int32),
        textures=fl.textures,
    )
    validate(flat)
    return flat


def flatten_camera(cam):
    rec = np.zeros((), dtype=CAMERA_DT)
    rec["look_from"] = _v3(cam.look_from)
    rec["right"], rec["up"], rec["fwd"] = _v3(cam.cameraRight), _v3(cam.cameraUp), _v3(cam.cameraFwd)
    rec["cam_w"], rec["cam_h"] = cam.camera_width, cam.camera_height
    rec["lens_radius"], rec["focal_distance"] = cam.lens_radius, cam.focal_distance
    rec["width"], rec["height"] = cam.screen_width, cam.screen_height
    return rec


def flatten_light(l):
    rec = np.zeros((), dtype=LIGHT_DT)
    if _kind(l) == "DirectionalLight":
        rec["kind"], rec["vec"] = LIGHT_DIRECTIONAL, _v3(l.Ldir)
    elif _kind(l) == "PointLight":
        rec["kind"], rec["vec"] = LIGHT_POINT, _v3(l.pos)
    else:
        raise TypeError(f"light {_kind(l)} is not supported by the CUDA backend")
    rec["color"] = _v3(l.color)
    return rec


def validate(flat):
    """Reject descriptions the reference itself cannot evaluate instead of
    silently inventing behaviour."""
    for ci in range(len(flat.colliders)):
        ctype = int(flat.colliders["type"][ci])
        mat = flat.materials[int(flat.primitives["material"][int(flat.colliders["primitive"][ci])])]
        uses_uv = (mat["color_tex"] >= 0 and mat["kind"] != MAT_SKYBOX) or mat["normalmap_tex"] >= 0 \
            or mat["kind"] in (MAT_THINFILM, MAT_SKYBOX)
        if ctype == COLLIDER_TRIANGLE and uses_uv:
            raise ValueError("triangles have no uv mapping (reference triangle.py:79-83 is broken): "
                             "use solid colours and no normal map on Triangle/TriangleMesh")
        if mat["normalmap_tex"] >= 0 and ctype not in (COLLIDER_PLANE, COLLIDER_CUBOID):
            raise ValueError("normal maps need a tangent frame: only Plane and Cuboid provide one "
                             "(reference material.py:32)")
    if len(flat.colliders) > 16382:
        raise ValueError("at most 16382 colliders per scene (ray records carry a 14-bit source id)")
    if len(flat.media) > 255:
        raise ValueError("at most 255 distinct refractive media per scene")


def round_to_float32(flat):
    """Copy of ``flat`` whose real parameters are rounded to float32 (what the device sees)."""
    import copy
    out = copy.copy(flat)
    def r(a):
        return a.astype(np.float32).astype(np.float64)
    out.ambient = r(flat.ambient)
    out.media = (flat.media.real.astype(np.float32) + 1j * flat.media.imag.astype(np.float32)).astype(np.complex128)
    for name in ("camera", "materials", "primitives", "colliders", "lights"):
        arr = getattr(flat, name).copy()
        for fname in arr.dtype.names:
            if arr.dtype[fname].base == np.float64:
                arr[fname] = r(arr[fname])
        setattr(out, name, arr)
    return out
