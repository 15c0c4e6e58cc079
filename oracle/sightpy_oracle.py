"""CPU oracle for the sightpy hot path — TEST INFRASTRUCTURE, NOT A PRODUCT PATH.

A float64 numpy restatement of what the reference (lmondada/Python-Raytracer) computes between
``Scene.render`` and the PIL image, written against the flattened POD scene description
(python-raytracer_b200/sightpy/flatten.py) that the CUDA library consumes as well.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may import this module; the package itself never does (it has no CPU fallback).

Pinned to the reference: ``tests/golden/make_golden.py`` imports the real reference from
``/root/reference`` (only possible in the build container), renders the same rays with it and
with this oracle under the same ``np.random.seed`` (``rng="legacy"`` consumes the numpy global
stream in the reference's order) and stores the reference outputs as fixtures;
``tests/test_oracle_golden.py`` replays them.  The reference itself ships no tests.

Two random-number modes:
  * ``legacy``: numpy global stream, reference draw order  -> bit-comparable to the reference;
  * ``philox``: the counter-based generator of the CUDA path, keyed by (pixel, path node)
                -> per-ray comparable to the GPU even for Monte-Carlo scenes.

Each function cites the reference lines it restates.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent.parent / "python-raytracer_b200"
if str(_PKG) not in sys.path:
    sys.path.insert(0, str(_PKG))

from sightpy.flatten import (  # noqa: E402
    COLLIDER_CUBOID, COLLIDER_PLANE, COLLIDER_SPHERE, COLLIDER_TRIANGLE, LIGHT_DIRECTIONAL, LIGHT_POINT,
    MAT_DIFFUSE, MAT_EMISSIVE, MAT_GLOSSY, MAT_REFRACTIVE, MAT_SKYBOX, MAT_THINFILM, FlatScene)

FARAWAY = 1.0e39          # constants.py:3
UPWARDS, UPDOWN = 1.0, -1.0
SKYBOX_DISTANCE = 1.0e6
NUDGE = 0.000001          # glossy.py:35, refractive.py:32/83, thin_film_interference.py:77/97, diffuse.py:36/91
WAVELENGTHS = np.array([630.0, 550.0, 475.0])   # refractive.py:113-121


# =================================================================================================
# small vector helpers on (3, N) arrays — same operation order as vector3.py so float64 results
# agree with the reference to the last bit wherever that is cheap to guarantee
# =================================================================================================
def dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def cross(a, b):
    return np.stack([a[1] * b[2] - a[2] * b[1], -a[0] * b[2] + a[2] * b[0], a[0] * b[1] - a[1] * b[0]])


def normalize(a):
    mag = np.sqrt(dot(a, a))
    return a * (1.0 / np.where(mag == 0, 1, mag))


def col(v):
    """(3,) -> (3, 1) so that it broadcasts against (3, N)."""
    return np.asarray(v).reshape(3, 1)


def matvec(m, a):
    """3x3 matrix times every column of (3, N) (vector3.py:93-97)."""
    return np.tensordot(np.asarray(m).reshape(3, 3), a, axes=([1], [0]))


# =================================================================================================
# Philox4x32-10 + path hashing (mirrors csrc/sp_rng.cuh bit for bit)
# =================================================================================================
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32).copy() for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(rounds):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & _MASK32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & _MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0, k1 = np.uint32(k0 + _W0), np.uint32(k1 + _W1)
    return c0, c1, c2, c3


def u01(word):
    """24-bit uniform in [0, 1): exactly representable in float32 and float64."""
    return (word >> np.uint32(8)).astype(np.float64) * (1.0 / 16777216.0)


def child_path(path, k):
    """Hash of the k-th child of a path-tree node (sp_rng.cuh: sp_child_path)."""
    with np.errstate(over="ignore"):
        x = (np.asarray(path, dtype=np.uint32) * np.uint32(0x01000193)) ^ \
            ((np.asarray(k, dtype=np.uint32) + np.uint32(1)) * np.uint32(0x9E3779B9))
        x ^= x >> np.uint32(16)
        x = x * np.uint32(0x85EBCA6B)
        x ^= x >> np.uint32(13)
        x = x * np.uint32(0xC2B2AE35)
        x ^= x >> np.uint32(16)
    return x.astype(np.uint32)


def root_path(sample):
    return child_path(np.uint32(0x811C9DC5), sample)


BLOCK_DIRECTION, BLOCK_MATERIAL = 0, 1


class PhiloxRng:
    mode = "philox"

    def __init__(self, seed=0):
        self.k0, self.k1 = np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)

    def draw4(self, pix, path, block):
        w = philox4x32(pix, path, np.uint32(block), np.uint32(0), self.k0, self.k1)
        return [u01(x) for x in w]


class LegacyRng:
    """numpy global stream, consumed in the reference's order."""
    mode = "legacy"

    @staticmethod
    def rand(n):
        return np.random.rand(n)


# =================================================================================================
class Bundle:
    """A batch of rays sharing depth / diffuse-bounce counters, as in ray.py:7-33."""
    __slots__ = ("O", "D", "medium", "depth", "dr", "pix", "path")

    def __init__(self, O, D, medium, depth, dr, pix, path):
        self.O, self.D, self.medium, self.depth, self.dr, self.pix, self.path = O, D, medium, depth, dr, pix, path

    def __len__(self):
        return self.O.shape[1]

    def take(self, mask):
        return Bundle(self.O[:, mask], self.D[:, mask], self.medium[mask], self.depth, self.dr,
                      self.pix[mask], self.path[mask])


class Oracle:
    def __init__(self, flat: FlatScene, rng="philox", seed=0):
        self.f = flat
        self.rng = PhiloxRng(seed) if rng == "philox" else LegacyRng()
        self._tex = {}
        self.rays_per_depth = {}
        self.shadow_rays = 0
        self.max_levels = 0          # debugging aid: rays of depth >= max_levels return black (0 = off)
        C = flat.colliders
        self.ctype = [int(t) for t in C["type"]]
        self.cprim = [int(p) for p in C["primitive"]]
        self.payload = C["p"]
        self.shadow_ids = [int(i) for i in flat.shadow_colliders]

    # ---- bookkeeping ---------------------------------------------------------------------------
    def tex(self, tid):
        if tid not in self._tex:
            self._tex[tid] = self.f.textures[tid].as_float()
        return self._tex[tid]

    def slot(self, ci, name):
        return self.f.field_of(ci, name)

    def reset_counters(self):
        self.rays_per_depth, self.shadow_rays = {}, 0

    @property
    def rays_total(self):
        return sum(self.rays_per_depth.values())

    # =============================================================================================
    # camera  (camera.py:51-85)
    # =============================================================================================
    def pixel_grid(self):
        cam = self.f.camera
        W, H = int(cam["width"]), int(cam["height"])
        x = np.linspace(-cam["cam_w"] / 2.0, cam["cam_w"] / 2.0, W)
        y = np.linspace(cam["cam_h"] / 2.0, -cam["cam_h"] / 2.0, H)
        xx, yy = np.meshgrid(x, y)
        return xx.flatten(), yy.flatten()

    def camera_rays(self, sample=0, pixels=None):
        """Jittered primary rays of one sample -> (O, D) as (3, N) float64 (+ pixel ids)."""
        cam = self.f.camera
        W, H = int(cam["width"]), int(cam["height"])
        gx, gy = self.pixel_grid()
        pix = np.arange(W * H, dtype=np.uint32) if pixels is None else np.asarray(pixels, dtype=np.uint32)
        gx, gy = gx[pix], gy[pix]
        n = len(pix)
        if self.rng.mode == "legacy":
            jx, jy = self.rng.rand(n), self.rng.rand(n)
            lr, lphi = self.rng.rand(n), self.rng.rand(n)            # random_in_unit_disk, random.py:6-9
        else:
            jx, jy, lr, lphi = self.rng.draw4(pix, np.broadcast_to(root_path(np.uint32(sample)), pix.shape),
                                              BLOCK_DIRECTION)
        x = gx + (jx - 0.5) * cam["cam_w"] / W
        y = gy + (jy - 0.5) * cam["cam_h"] / H
        r, phi = np.sqrt(lr), lphi * 2 * np.pi
        rx, ry = r * np.cos(phi), r * np.sin(phi)
        lf, right, up, fwd = (col(cam[k]) for k in ("look_from", "right", "up", "fwd"))
        lens, fd = cam["lens_radius"], cam["focal_distance"]
        origin = lf + right * rx * lens + up * ry * lens
        direction = normalize(lf + up * y * fd + right * x * fd + fwd * fd - origin)
        return origin, direction, pix

    # =============================================================================================
    # colliders: intersect -> (distance, orientation), FARAWAY on miss
    # =============================================================================================
    def intersect(self, ci, O, D):
        return (self._sphere, self._plane, self._cuboid, self._triangle)[self.ctype[ci]](ci, O, D)

    def _sphere(self, ci, O, D):                                    # sphere.py:26-52
        C, r = col(self.slot(ci, "center")), self.slot(ci, "radius")
        b = 2 * dot(D, O - C)
        c = dot(C, C) + dot(O, O) - 2 * dot(C, O) - (r * r)
        disc = (b ** 2) - (4 * c)
        sq = np.sqrt(np.maximum(0, disc))
        h0, h1 = (-b - sq) / 2, (-b + sq) / 2
        h = np.where((h0 > 0) & (h0 < h1), h0, h1)
        pred = (disc > 0) & (h > 0)
        NdotD = dot((O + D * h - C) * (1.0 / r), D)
        t = np.where(pred & (NdotD != 0), h, FARAWAY)
        orient = np.where(pred & (NdotD > 0), UPDOWN, np.where(pred & (NdotD < 0), UPWARDS, FARAWAY))
        return t, orient

    def _planar(self, N, C, O, D):
        """Shared ray/plane step of plane.py:57-67 and triangle.py:37-46."""
        NdotD = dot(N, D)
        NdotD = np.where(NdotD == 0.0, NdotD + 0.0001, NdotD)
        NdotC_O = dot(N, C - O)
        d = D * NdotC_O / NdotD
        M = O + d
        return NdotD, NdotC_O, M, np.sqrt(dot(d, d))

    def _plane(self, ci, O, D):                                     # plane.py:57-90
        N, C = col(self.slot(ci, "normal")), col(self.slot(ci, "center"))
        NdotD, k, M, dis = self._planar(N, C, O, D)
        M_C = M - C
        u, v = dot(col(self.slot(ci, "u_axis")), M_C), dot(col(self.slot(ci, "v_axis")), M_C)
        inside = (np.abs(u) <= self.slot(ci, "w")) & (np.abs(v) <= self.slot(ci, "h")) & (k * NdotD > 0)
        return np.where(inside, dis, FARAWAY), np.where(inside, np.where(NdotD < 0, UPWARDS, UPDOWN), FARAWAY)

    def _triangle(self, ci, O, D):                                  # triangle.py:37-66
        N, C = col(self.slot(ci, "normal")), col(self.slot(ci, "centroid"))
        NdotD, k, M, dis = self._planar(N, C, O, D)
        inside = ((dot(col(self.slot(ci, "n31")), M - col(self.slot(ci, "p1"))) >= 0)
                  & (dot(col(self.slot(ci, "n12")), M - col(self.slot(ci, "p2"))) >= 0)
                  & (dot(col(self.slot(ci, "n23")), M - col(self.slot(ci, "p3"))) >= 0)
                  & (k * NdotD > 0))
        return np.where(inside, dis, FARAWAY), np.where(inside, np.where(NdotD < 0, UPWARDS, UPDOWN), FARAWAY)

    def _cuboid(self, ci, O, D):                                    # cuboid.py:105-140
        B = self.slot(ci, "basis")
        Ol, Dl = matvec(B, O), matvec(B, D)
        lb, rt = self.slot(ci, "lb_local"), self.slot(ci, "rt_local")
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / Dl
            t1, t2 = (lb[0] - Ol[0]) * inv[0], (rt[0] - Ol[0]) * inv[0]
            t3, t4 = (lb[1] - Ol[1]) * inv[1], (rt[1] - Ol[1]) * inv[1]
            t5, t6 = (lb[2] - Ol[2]) * inv[2], (rt[2] - Ol[2]) * inv[2]
            tmin = np.maximum(np.maximum(np.minimum(t1, t2), np.minimum(t3, t4)), np.minimum(t5, t6))
            tmax = np.minimum(np.minimum(np.maximum(t1, t2), np.maximum(t3, t4)), np.maximum(t5, t6))
            miss = (tmax < 0) | (tmin > tmax)
            inside = tmin < 0
        t = np.where(miss, FARAWAY, np.where(inside, tmax, tmin))
        orient = np.where(miss, FARAWAY, np.where(inside, UPDOWN, UPWARDS))
        return t, orient

    # ---- geometric normal / uv -------------------------------------------------------------------
    def collider_normal(self, ci, P):
        ct = self.ctype[ci]
        if ct == COLLIDER_SPHERE:                                   # sphere.py:54-56
            return (P - col(self.slot(ci, "center"))) * (1.0 / self.slot(ci, "radius"))
        if ct in (COLLIDER_PLANE, COLLIDER_TRIANGLE):               # plane.py:104-105, triangle.py:85-86
            return np.broadcast_to(col(self.slot(ci, "normal")), P.shape).copy()
        size = self.slot(ci, "size")                                # cuboid.py:142-151
        Pl = matvec(self.slot(ci, "basis"), P - col(self.slot(ci, "center")))
        a = col([1.0 / size[0], 1.0 / size[1], 1.0 / size[2]]) * np.abs(Pl)
        amax = np.maximum(np.maximum(a[0], a[1]), a[2])
        face = np.stack([np.where(amax == a[i], np.sign(Pl[i]), 0.0) for i in range(3)])
        return matvec(self.slot(ci, "inv_basis"), face)

    def collider_uv(self, ci, P):
        ct = self.ctype[ci]
        if ct == COLLIDER_SPHERE:                                   # sphere.py:58-64
            m = (P - col(self.slot(ci, "center"))) / self.slot(ci, "radius")
            with np.errstate(invalid="ignore"):
                return (np.arctan2(m[2], m[0]) + np.pi) / (2 * np.pi), (np.arcsin(m[1]) + np.pi / 2) / np.pi
        if ct == COLLIDER_PLANE:                                    # plane.py:98-102
            M_C = P - col(self.slot(ci, "center"))
            sh = self.slot(ci, "uv_shift")
            return ((dot(col(self.slot(ci, "u_axis")), M_C) / self.slot(ci, "w") + 1) / 2 + sh[0],
                    (dot(col(self.slot(ci, "v_axis")), M_C) / self.slot(ci, "h") + 1) / 2 + sh[1])
        if ct == COLLIDER_CUBOID:                                   # cuboid.py:153-187
            N = self.collider_normal(ci, P)
            M_C = P - col(self.slot(ci, "center"))
            aw, ah, al = (col(self.slot(ci, k)) for k in ("ax_w", "ax_h", "ax_l"))
            width = self.slot(ci, "size")[0]

            def is_face(x, y, z):
                return (N[0] == x) & (N[1] == y) & (N[2] == z)

            faces = [is_face(0., -1., 0.), is_face(0., 1., 0.), is_face(1., 0., 0.),
                     is_face(-1., 0., 0.), is_face(0., 0., 1.), is_face(0., 0., -1.)]

            def g(axis, off):
                return (dot(axis, M_C) / width * 2 * 0.985 + 1) / 2 + off

            u = np.select(faces, [g(aw, 1), g(aw, 1), g(al, 2), g(al * -1, 0), g(aw * -1, 3), g(aw, 1)])
            v = np.select(faces, [g(al * -1, 0), g(al, 2), g(ah, 1), g(ah, 1), g(ah, 1), g(ah, 1)])
            return u, v
        raise ValueError("triangles have no uv mapping (triangle.py:79-83 is broken upstream)")

    def hit_uv(self, ci, P):
        u, v = self.collider_uv(ci, P)
        if self.f.primitives["uv_cross_layout"][self.cprim[ci]]:    # cuboid.py:29-32, skybox.py:29-32
            u, v = u / 4, v / 3
        return u, v

    @staticmethod
    def texel_index(u, v, H, W, repeat):
        """Python negative-index row / wrapped column of texture.py:34-37 and its clones."""
        with np.errstate(invalid="ignore"):
            row = -((v * H * repeat).astype(int) % H)
            colm = (u * W * repeat).astype(int) % W
        return row, colm

    def sample_texture(self, tid, u, v, repeat, index_shape=None):
        img = self.tex(tid)
        H, W = index_shape if index_shape is not None else img.shape[:2]
        row, colm = self.texel_index(u, v, H, W, repeat)
        return img[row, colm].T                                      # (3, N)

    def shading_normal(self, ci, mat, P, orient):                   # material.py:18-36
        if mat["normalmap_tex"] >= 0:
            u, v = self.hit_uv(ci, P)
            im = self.sample_texture(int(mat["normalmap_tex"]), u, v, mat["normalmap_repeat"])
            n_map = (im - 0.5) * 2.0
            return normalize(matvec(self.slot(ci, "inv_basis"), n_map)) * orient
        return self.collider_normal(ci, P) * orient

    def material_color(self, ci, mat, P):
        """solid_color / image lookup (texture.py:23-39)."""
        if mat["color_tex"] < 0:
            return np.broadcast_to(col(mat["color"]), P.shape)
        u, v = self.hit_uv(ci, P)
        return self.sample_texture(int(mat["color_tex"]), u, v, mat["color_repeat"])

    # =============================================================================================
    # integrator  (ray.py:122-148)
    # =============================================================================================
    def radiance(self, b: Bundle, want_hits=False):
        n = len(b)
        if self.max_levels and b.depth >= self.max_levels and not want_hits:
            return np.zeros((3, n))
        self.rays_per_depth[b.depth] = self.rays_per_depth.get(b.depth, 0) + n
        inters = [self.intersect(ci, b.O, b.D) for ci in range(len(self.ctype))]
        nearest = inters[0][0]
        for t, _ in inters[1:]:
            nearest = np.minimum(nearest, t)
        color = np.zeros((3, n))
        hit_id = np.full(n, -1, dtype=np.int32)
        for ci, (t, orient) in enumerate(inters):
            mask = (nearest != FARAWAY) & (t == nearest)
            if np.any(mask):
                hit_id[mask & (hit_id < 0)] = ci
                cc = self.shade(ci, b.take(mask), t[mask], orient[mask])
                color[:, mask] += cc
        if want_hits:
            return color, hit_id, np.where(nearest == FARAWAY, np.inf, nearest)
        return color

    def shade(self, ci, b, t, orient):
        prim = self.f.primitives[self.cprim[ci]]
        mat = self.f.materials[int(prim["material"])]
        kind = int(mat["kind"])
        P = b.O + b.D * t
        if kind == MAT_EMISSIVE:                                     # emissive.py:21-23
            return np.array(self.material_color(ci, mat, P), dtype=np.float64)
        if kind == MAT_SKYBOX:
            return self.shade_skybox(ci, mat, b, P)
        if kind == MAT_GLOSSY:
            return self.shade_glossy(ci, prim, mat, b, P, orient)
        if kind == MAT_REFRACTIVE:
            return self.shade_refractive(ci, prim, mat, b, P, t, orient)
        if kind == MAT_THINFILM:
            return self.shade_thinfilm(ci, prim, mat, b, P, orient)
        if kind == MAT_DIFFUSE:
            return self.shade_diffuse(ci, prim, mat, b, P, orient)
        raise ValueError(f"unknown material kind {kind}")

    def child(self, b, O, D, medium, k, dr=None, mask=None):
        """Bundle of secondary rays one level deeper; k = child slot in the path tree."""
        path = child_path(b.path, k)
        pix = b.pix
        if mask is not None:
            O, D, medium, path, pix = O[:, mask], D[:, mask], medium[mask], path[mask], pix[mask]
        return Bundle(O, D, medium, b.depth + 1, b.dr if dr is None else dr, pix, path)

    # ---- SkyBox / Panorama  (skybox.py:51-94) ----------------------------------------------------
    def shade_skybox(self, ci, mat, b, P):
        u, v = self.hit_uv(ci, P)
        color = np.array(self.sample_texture(int(mat["color_tex"]), u, v, mat["color_repeat"]), dtype=np.float64)
        if b.depth != 0 and mat["light_intensity"] != 0.0:
            ls = self.sample_texture(int(mat["aux_tex0"]), u, v, mat["color_repeat"],
                                     index_shape=(int(mat["index_h"]), int(mat["index_w"])))
            color = color + mat["light_intensity"] * ls
        return color

    # ---- Glossy  (glossy.py:25-110, lights.py:25-52) ---------------------------------------------
    def shade_glossy(self, ci, prim, mat, b, P, orient):
        f = self.f
        N = self.shading_normal(ci, mat, P, orient)
        diff = self.material_color(ci, mat, P) * mat["diff_coeff"]
        color = col(f.ambient) * diff
        V = b.D * -1.0
        nudged = P + N * NUDGE
        n_mat = col(mat["n_re"] + 1j * mat["n_im"])
        n_ray = f.media[b.medium].T                                   # (3, N) complex
        for light in f.lights:
            if light["kind"] == LIGHT_DIRECTIONAL:
                L = np.broadcast_to(col(light["vec"]), P.shape)
                dist = SKYBOX_DISTANCE
                NdotL = np.maximum(dot(N, L), 0.0)
                lv = col(light["color"]) * NdotL
            elif light["kind"] == LIGHT_POINT:                       # intended behaviour of lights.py:25-37
                to_l = col(light["vec"]) - P
                dist = np.sqrt(dot(to_l, to_l))
                L = to_l * (1.0 / dist)
                NdotL = np.maximum(dot(N, L), 0.0)
                lv = col(light["color"]) * NdotL / (dist ** 2.0) * 100
            H = normalize(L + V)
            if self.shadow_ids:
                self.shadow_rays += P.shape[1]
                near = None
                for si in self.shadow_ids:
                    ts, _ = self.intersect(si, nudged, L)
                    near = ts if near is None else np.minimum(near, ts)
                see = (near >= dist).astype(np.float64)
            else:
                see = 1.0
            color = color + diff * lv * see
            if mat["roughness"] != 0.0:
                F0 = np.abs((n_ray - n_mat) / (n_ray + n_mat)) ** 2
                cos_t = np.clip(dot(V, H), 0.0, 1.0)
                F = F0 + (1.0 - F0) * (1.0 - cos_t) ** 5
                a = 2.0 / (mat["roughness"] ** 2.0) - 2.0
                Dphong = np.power(np.clip(dot(N, H), 0.0, 1.0), a) * (a + 2.0) / (2.0 * np.pi)
                color = color + (F * Dphong / (4.0 * np.clip(dot(N, V) * NdotL, 0.001, 1.0))
                                 * see * lv * mat["spec_coeff"])
        if b.depth < prim["max_ray_depth"]:
            n_scene = col(f.media[0])
            F0 = np.abs((n_scene - n_mat) / (n_scene + n_mat)) ** 2
            cos_t = np.clip(dot(V, N), 0.0, 1.0)
            F = F0 + (1.0 - F0) * (1.0 - cos_t) ** 5
            R = normalize(b.D - N * 2.0 * dot(b.D, N))
            color = color + self.radiance(self.child(b, nudged, R, b.medium, 0)) * F
        return color

    # ---- Refractive  (refractive.py:24-123) --------------------------------------------------------
    def shade_refractive(self, ci, prim, mat, b, P, t, orient):
        f = self.f
        n = P.shape[1]
        color = np.zeros((3, n))
        if not (b.depth < prim["max_ray_depth"]):
            return color
        N = self.shading_normal(ci, mat, P, orient)
        V = b.D * -1.0
        nudged = P + N * NUDGE
        n1 = f.media[b.medium].T
        med2 = np.where(orient == UPWARDS, int(mat["medium"]), 0)
        n2 = f.media[med2].T
        n1_div_n2 = n1.real / n2.real
        cos_i = dot(V, N)
        cos_t = np.sqrt(1.0 - (n1 / n2) ** 2 * (1.0 - cos_i ** 2))
        r_per = (n1 * cos_i - n2 * cos_t) / (n1 * cos_i + n2 * cos_t)
        r_par = -1.0 * (n1 * cos_t - n2 * cos_i) / (n1 * cos_t + n2 * cos_i)
        F = (np.abs(r_per) ** 2 + np.abs(r_par) ** 2) / 2.0
        T = 1.0 - F
        R = normalize(b.D - N * 2.0 * dot(b.D, N))
        eta = (n1_div_n2[0] + n1_div_n2[1] + n1_div_n2[2]) / 3
        sin2_t = eta ** 2 * (1.0 - cos_i ** 2)
        non_tir = sin2_t <= 1.0
        refr_dir = normalize(b.D * eta + N * (eta * cos_i - np.sqrt(1 - np.clip(sin2_t, 0, 1))))
        nudged_in = P - N * NUDGE
        if prim["mc"]:
            if self.rng.mode == "legacy":
                xi = self.rng.rand(n)
            else:
                xi = self.rng.draw4(b.pix, b.path, BLOCK_MATERIAL)[0]
            pick = (xi > (F[0] + F[1] + F[2]) / 3) & non_tir
            child = self.child(b, np.where(pick, nudged_in, nudged), np.where(pick, refr_dir, R),
                               np.where(pick, med2, b.medium), 0)
            color = self.radiance(child)
        else:
            color = self.radiance(self.child(b, nudged, R, b.medium, 0)) * F
            if np.any(non_tir):
                sub = self.radiance(self.child(b, nudged_in, refr_dir, med2, 1, mask=non_tir))
                placed = np.zeros((3, n))
                placed[:, non_tir] = sub
                color = color + placed * T
        absorb = np.exp(-2.0 * n1.imag * 2.0 * np.pi / col(WAVELENGTHS) * 1e9 * t)
        return color * absorb

    # ---- ThinFilmInterference  (thin_film_interference.py:24-115) ----------------------------------
    def shade_thinfilm(self, ci, prim, mat, b, P, orient):
        n = P.shape[1]
        color = np.zeros((3, n))
        if not (b.depth < prim["max_ray_depth"]):
            return color
        N = self.shading_normal(ci, mat, P, orient)
        V = b.D * -1.0
        cos_i = dot(V, N)
        u, v = self.hit_uv(ci, P)
        lut = self.tex(int(mat["aux_tex0"]))
        if mat["noise_factor"] != 0.0:
            noise = self.tex(int(mat["aux_tex1"]))[:, :, 0]
            row, colm = self.texel_index(u, v, noise.shape[0], noise.shape[1], 0.5)
            thick = mat["thickness"] + mat["noise_factor"] * (noise[row, colm] - 0.5)
            Fim = lut[(cos_i * lut.shape[0]).astype(int), thick.astype(int)]
        else:
            Fim = lut[(cos_i * lut.shape[0]).astype(int), int(mat["thickness"])]
        F = Fim.T
        R = normalize(b.D - N * 2.0 * dot(b.D, N))
        color = color + (col(self.f.ambient) + self.radiance(self.child(b, P + N * NUDGE, R, b.medium, 0))) * F
        color = color + self.radiance(self.child(b, P - N * NUDGE, b.D, b.medium, 1)) * (1.0 - F)
        return color

    # ---- Diffuse + sampling pdfs  (diffuse.py:25-124, random.py:50-174) ------------------------------
    @staticmethod
    def _onb(w):
        """Tangent frame of cosine_pdf / spherical_caps_pdf (random.py:60-63, 112-115)."""
        a = np.where(np.abs(w[0]) > 0.9, col([0, 1, 0]), col([1, 0, 0])).astype(np.float64)
        v = normalize(cross(w, a))
        return cross(w, v), v

    def sample_diffuse(self, N, origin, pix, path, weight_cos):
        """Direction + pdf value of the cosine / spherical-cap mixture for every ray."""
        f = self.f
        n = N.shape[1]
        imp = [f.primitives[int(i)] for i in f.importance]
        l = len(imp)
        legacy = self.rng.mode == "legacy"
        if legacy:
            xi_mix = self.rng.rand(n) if l else None
            phi_c, r2_c = self.rng.rand(n) * 2 * np.pi, self.rng.rand(n)
        else:
            xi_mix, u_phi, u_r2, xi_pick = self.rng.draw4(pix, path, BLOCK_DIRECTION)
            phi_c, r2_c = u_phi * 2 * np.pi, u_r2
        cu, cv = self._onb(N)
        d_cos = cu * (np.cos(phi_c) * np.sqrt(r2_c)) + cv * (np.sin(phi_c) * np.sqrt(r2_c)) + N * np.sqrt(1 - r2_c)
        if l == 0:
            return d_cos, np.clip(dot(d_cos, N), 0.0, 1.0) / np.pi
        # spherical caps towards the importance-sampled primitives (random.py:96-150)
        if legacy:
            pick = (self.rng.rand(n) * l).astype(int)
            phi_s, r2_s = self.rng.rand(n) * 2 * np.pi, self.rng.rand(n)
        else:
            pick = (xi_pick * l).astype(int)
            phi_s, r2_s = phi_c, r2_c
        ws, cmaxs, us, vs = [], [], [], []
        for p in imp:
            to_c = col(p["center"]) - origin
            w = normalize(to_c)
            uu, vv = self._onb(w)
            dist = np.sqrt(dot(to_c, to_c))
            with np.errstate(divide="ignore", invalid="ignore"):
                cmax = np.sqrt(1 - np.clip(p["bounded_sphere_radius"] / dist, 0.0, 1.0) ** 2)
            ws.append(w); cmaxs.append(cmax); us.append(uu); vs.append(vv)
        masks = [pick == i for i in range(l)]
        cmax = np.select(masks, cmaxs)
        w = np.stack([np.select(masks, [x[k] for x in ws]) for k in range(3)])
        vv = np.stack([np.select(masks, [x[k] for x in vs]) for k in range(3)])
        uu = np.stack([np.select(masks, [x[k] for x in us]) for k in range(3)])
        z = 1.0 + r2_s * (cmax - 1.0)
        s = np.sqrt(1.0 - z ** 2)
        d_cap = uu * (np.cos(phi_s) * s) + vv * (np.sin(phi_s) * s) + w * z
        direction = np.where(xi_mix < weight_cos, d_cos, d_cap)      # mixed_pdf.generate, random.py:170-174
        pdf_cos = np.clip(dot(direction, N), 0.0, 1.0) / np.pi
        pdf_cap = 0.0
        with np.errstate(divide="ignore", invalid="ignore"):
            for wi, ci_ in zip(ws, cmaxs):
                pdf_cap = pdf_cap + np.where(dot(direction, wi) > ci_, 1 / ((1 - ci_) * 2 * np.pi), 0.0)
        pdf_cap = pdf_cap / l
        return direction, pdf_cos * weight_cos + pdf_cap * (1.0 - weight_cos)

    def shade_diffuse(self, ci, prim, mat, b, P, orient):
        n = P.shape[1]
        N = self.shading_normal(ci, mat, P, orient)
        diff = self.material_color(ci, mat, P)
        nudged = P + N * NUDGE
        if b.dr < 1:
            m = int(mat["diffuse_rays"])
            N_r, O_r = np.repeat(N, m, axis=1), np.repeat(nudged, m, axis=1)
            pix_r, med_r = np.repeat(b.pix, m), np.repeat(b.medium, m)
            path_r = child_path(np.repeat(b.path, m), np.tile(np.arange(m, dtype=np.uint32), n))
            d, pdf = self.sample_diffuse(N_r, O_r, pix_r, path_r, mat["ambient_weight"])
            NdotL = np.clip(dot(d, N_r), 0.0, 1.0)
            kids = Bundle(O_r, d, med_r, b.depth + 1, b.dr + 1, pix_r, path_r)
            with np.errstate(divide="ignore", invalid="ignore"):
                Lk = self.radiance(kids) * NdotL / pdf / np.pi
            return diff * Lk.reshape(3, n, m).mean(axis=2)
        if b.dr < mat["max_diffuse_reflections"]:
            path = child_path(b.path, 0)
            d, pdf = self.sample_diffuse(N, nudged, b.pix, path, mat["ambient_weight"])
            NdotL = np.clip(dot(N, d), 0.0, 1.0)
            kids = Bundle(nudged, d, b.medium, b.depth + 1, b.dr + 1, b.pix, path)
            with np.errstate(divide="ignore", invalid="ignore"):
                return diff * self.radiance(kids) * NdotL / pdf / np.pi
        return np.zeros((3, n))

    # =============================================================================================
    # entry points
    # =============================================================================================
    def trace_columns(self, O, D, pix=None, sample=0):
        """get_raycolor on caller rays given as (3, N) float64 columns."""
        n = O.shape[1]
        pix = np.arange(n, dtype=np.uint32) if pix is None else np.asarray(pix, dtype=np.uint32)
        b = Bundle(O, D, np.zeros(n, dtype=np.int64), 0, 0, pix,
                   np.broadcast_to(root_path(np.uint32(sample)), (n,)).copy())
        color, hit, t = self.radiance(b, want_hits=True)
        return dict(rgb=color.T.copy(), hit_id=hit, t=t)

    def trace(self, origins, dirs, pix=None, sample=0, renormalize=True):
        """Same with (N, 3) arrays (the sp_trace layout) -> dict(rgb (N,3), hit_id (N,), t (N,)).

        Directions are renormalised in float64 first.  The reference only ever traces exactly
        normalised directions (camera.py:85) and its hit points are ``O + D * |D t|``
        (plane.py:66-67), so a float32-rounded direction (|D| = 1 +- 3e-8) would move every hit
        point by t * 3e-8 along the ray — more than the 1e-6 origin nudge of its secondary rays
        once t > 30, which makes children start *behind* the surface they leave.  That is an
        artefact of handing float32 test rays to float64 code, not reference behaviour."""
        O = np.ascontiguousarray(np.asarray(origins, dtype=np.float64).T)
        D = np.ascontiguousarray(np.asarray(dirs, dtype=np.float64).T)
        if renormalize:
            D = D / np.sqrt(dot(D, D))
        return self.trace_columns(O, D, pix, sample)

    def nearest(self, origins, dirs):
        """Nearest collider index (-1 = none) and distance (inf = none) of (N, 3) rays: the first half of
        get_raycolor (ray.py:124-128) without any shading — affordable on scenes of thousands of colliders."""
        O = np.ascontiguousarray(np.asarray(origins, dtype=np.float64).T)
        D = np.ascontiguousarray(np.asarray(dirs, dtype=np.float64).T)
        D = D / np.sqrt(dot(D, D))
        best = np.full(O.shape[1], FARAWAY)
        hit = np.full(O.shape[1], -1, dtype=np.int32)
        for ci in range(len(self.ctype)):
            t, _ = self.intersect(ci, O, D)
            closer = t < best                         # ties: the lowest index is reported, as in radiance()
            hit[closer] = ci
            best = np.where(closer, t, best)
        return hit, np.where(best == FARAWAY, np.inf, best)

    def render_linear(self, spp, sample_begin=0):
        """Sum over samples / spp of get_raycolor(camera rays) -> (3, H*W) (scene.py:78-119)."""
        total = None
        for s in range(sample_begin, sample_begin + spp):
            O, D, pix = self.camera_rays(s)
            c = self.trace_columns(O, D, pix, sample=s)["rgb"].T
            total = c if total is None else total + c
        return total / spp

    def distances(self, sample=0):
        O, D, _ = self.camera_rays(sample)
        near = None
        for ci in range(len(self.ctype)):
            t, _ = self.intersect(ci, O, D)
            near = t if near is None else np.minimum(near, t)
        return np.where(near == FARAWAY, np.inf, near)


# =================================================================================================
# frame resolve  (colour_functions.py:4-18, scene.py:125-140)
# =================================================================================================
def tonemap_u8(linear_3xN, height, width):
    lin = np.asarray(linear_3xN, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        enc = np.where(lin <= 0.00304, 12.92 * lin, 1.055 * np.power(lin, 1.0 / 2.4) - 0.055)
    peak = np.amax(enc, axis=0) + 0.00001
    enc = np.where(peak > 1.0, enc * 1.0 / peak, enc)
    planes = [(255 * np.clip(c, 0, 1).reshape(height, width)).astype(np.uint8) for c in enc]
    return np.stack(planes, axis=2)
