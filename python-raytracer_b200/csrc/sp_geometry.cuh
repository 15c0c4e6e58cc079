// Ray / collider intersection over one staged geometry chunk (FP32, brute force, nearest hit).
//
// Restates Sphere_Collider.intersect (sphere.py:26-52), Plane_Collider.intersect (plane.py:57-90),
// Cuboid_Collider.intersect (cuboid.py:105-140) and Triangle_Collider.intersect (triangle.py:37-66)
// in forms that are stable in float32:
//   * sphere: discriminant from the perpendicular offset  r^2 - |oc - (D.oc) D|^2  instead of the
//     expanded  |C|^2 + |O|^2 - 2 C.O - r^2  (which cancels catastrophically in float32);
//   * plane / triangle: everything relative to O - C;  cuboid: slabs on B (O - C);
//   * "FARAWAY" (1e39) does not exist in float32: a miss is t = +inf.
// All lanes of a warp walk the same chunk in the same order, so every shared-memory read is a
// broadcast (no bank conflicts) and there is no divergence in the loop structure.
//
// Self-intersection: the reference offsets secondary-ray origins by 1e-6 along the normal, which
// is below float32 resolution at scene scale (ulp(555) = 6e-5).  Instead a ray record names the
// collider it starts on and how it leaves it (SP_SELF_*, sp_types.cuh) and that one collider is
// answered analytically — the eps -> 0 limit of what the reference computes in float64.
#pragma once
#include "sp_types.cuh"

struct HitRec {
    float t;        // parametric distance (directions are unit length), +inf = miss
    int id;         // index into scene.collider_list, -1 = miss
    int orient;     // +1 UPWARDS (outer face), -1 UPDOWN (inner face)
};

// where the ray's source collider sits inside the current chunk (-1 = not in this chunk)
struct SelfSlot { int sphere, plane, cuboid, tri, aa; uint32_t mode; };

struct ChunkBest { float t; int idx; int orient; };   // idx = position in the chunk's id array

// ---- one collider against one ray --------------------------------------------------------------------
// `tag` is what a winning test leaves in best.idx (position in the chunk's id array, or the collider id for
// BVH leaves); `is_self` marks the collider the ray starts on (see the header comment).
SP_DEV void sp_item_sphere(float4 s, float3 O, float3 D, bool is_self, uint32_t mode, int tag, ChunkBest& best) {
    float3 oc = O - xyz(s);
    float b = dot(D, oc);
    float3 q = fma3(D, -b, oc);
    float disc = s.w - dot(q, q);
    if (disc > 0.f) {
        float sq = fast_sqrt(disc);
        float h0 = -b - sq, h1 = -b + sq;
        bool near_ok = (h0 > 0.f) && !is_self;             // SP_SELF_FAR: only the far root
        float t = near_ok ? h0 : h1;
        bool ok = (t > 0.f) && !(is_self && mode != SP_SELF_FAR);
        if (ok && t < best.t) { best.t = t; best.idx = tag; best.orient = near_ok ? 1 : -1; }
    }
}

SP_DEV void sp_item_plane(float4 a, float4 c, float4 u4, float4 v4, float3 O, float3 D, bool is_self, int tag,
                          ChunkBest& best) {
    float3 N = xyz(a), oc = O - xyz(c);
    float nd = dot(N, D);
    nd = (nd == 0.f) ? 1e-4f : nd;
    float k = -dot(N, oc);
    float t = __fdividef(k, nd);
    float u = fmaf(t, dot(xyz(u4), D), dot(xyz(u4), oc));
    float v = fmaf(t, dot(xyz(v4), D), dot(xyz(v4), oc));
    bool ok = (fabsf(u) <= a.w) && (fabsf(v) <= c.w) && (k * nd > 0.f) && !is_self;
    if (ok && t < best.t) { best.t = t; best.idx = tag; best.orient = nd < 0.f ? 1 : -1; }
}

SP_DEV void sp_item_cuboid(float4 r0, float4 r1, float4 r2, float4 c, float4 e, float3 O, float3 D, bool is_self,
                           uint32_t mode, int tag, ChunkBest& best) {
    float3 oc = O - xyz(c);
    float3 Ol = v3(dot(xyz(r0), oc), dot(xyz(r1), oc), dot(xyz(r2), oc));
    float3 Dl = v3(dot(xyz(r0), D), dot(xyz(r1), D), dot(xyz(r2), D));
    float ix = fast_rcp(Dl.x), iy = fast_rcp(Dl.y), iz = fast_rcp(Dl.z);   // +-inf for axis-parallel rays, as 1/0
    float t1 = (r0.w - Ol.x) * ix, t2 = (c.w - Ol.x) * ix;
    float t3 = (r1.w - Ol.y) * iy, t4 = (e.x - Ol.y) * iy;
    float t5 = (r2.w - Ol.z) * iz, t6 = (e.y - Ol.z) * iz;
    float tmin = fmaxf(fmaxf(fminf(t1, t2), fminf(t3, t4)), fminf(t5, t6));
    float tmax = fminf(fminf(fmaxf(t1, t2), fmaxf(t3, t4)), fmaxf(t5, t6));
    bool miss = (tmax < 0.f) || (tmin > tmax);
    bool inside = (tmin < 0.f) || is_self;                 // SP_SELF_FAR: exit point only
    float t = inside ? tmax : tmin;
    bool ok = !miss && !(is_self && mode != SP_SELF_FAR);
    if (ok && t < best.t) { best.t = t; best.idx = tag; best.orient = inside ? -1 : 1; }
}

// triangle.py:37-66 through the affine map to the unit triangle: the three edge tests
// n31.(M-p1) >= 0, n12.(M-p2) >= 0, n23.(M-p3) >= 0 are u >= 0, v >= 0, 1-u-v >= 0 of the hit point's
// barycentric coordinates, and the third row of the map is the plane normal (N.D, N.(O-p1)).
SP_DEV void sp_item_triangle(float4 m0, float4 m1, float4 m2, float3 O, float3 D, bool is_self, int tag, ChunkBest& best) {
    float nd = dot(xyz(m2), D);
    nd = (nd == 0.f) ? 1e-4f : nd;
    const float w0 = dot(xyz(m2), O) + m2.w;                 // N.(O - p1) = -k
    const float t = -w0 * fast_rcp(nd);
    const float u = fmaf(t, dot(xyz(m0), D), dot(xyz(m0), O) + m0.w);
    const float v = fmaf(t, dot(xyz(m1), D), dot(xyz(m1), O) + m1.w);
    const bool ok = (u >= 0.f) && (v >= 0.f) && (u + v <= 1.f) && (t > 0.f) && !is_self;
    if (ok && t < best.t) { best.t = t; best.idx = tag; best.orient = nd < 0.f ? 1 : -1; }
}

// Axis-aligned rectangle, normal along axis A (0, 1, 2), in-plane axes B < C.  1/0 = inf sends the hit point of
// an axis-parallel ray out of bounds (the reference substitutes N.D = 1e-4 there and misses all the same).
template <int A>
SP_DEV void sp_item_aa(float4 r0, float4 r1, float3 O, float3 D, float inv_da, bool is_self, int tag, ChunkBest& best) {
    const float oa = A == 0 ? O.x : (A == 1 ? O.y : O.z), da = A == 0 ? D.x : (A == 1 ? D.y : D.z);
    const float ob = A == 0 ? O.y : O.x, db = A == 0 ? D.y : D.x;
    const float oc = A == 2 ? O.y : O.z, dc = A == 2 ? D.y : D.z;
    const float ca = A == 0 ? r0.x : (A == 1 ? r0.y : r0.z);
    const float cb = A == 0 ? r0.y : r0.x, cc = A == 2 ? r0.y : r0.z;
    const float t = (ca - oa) * inv_da;                        // k / N.D with the normal's sign cancelled
    const float pb = fmaf(t, db, ob) - cb, pc = fmaf(t, dc, oc) - cc;
    const bool ok = (fabsf(pb) <= r1.x) && (fabsf(pc) <= r1.y) && (t > 0.f) && !is_self;
    if (ok && t < best.t) { best.t = t; best.idx = tag; best.orient = (r0.w * da < 0.f) ? 1 : -1; }
}

template <int A>
SP_DEV void sp_intersect_aa(const float4* __restrict__ aa, int first, int count, int id_base, float3 O, float3 D,
                            float inv_da, int self_aa, ChunkBest& best) {
#pragma unroll 1
    for (int i = first; i < first + count; ++i)
        sp_item_aa<A>(aa[2 * i], aa[2 * i + 1], O, D, inv_da, i == self_aa, id_base + i, best);
}

// Loops are not unrolled: scenes of a few colliders run the remainder code of an unrolled loop anyway, and the
// smaller kernel is 2-9 % faster on the example scenes (instruction cache).  Only long runs of spheres / triangles
// (the exhaustive multi-chunk walk of large scenes, option "bvh" = 0) go through a 4-way unrolled bulk loop.
SP_DEV void sp_intersect_chunk(const float4* __restrict__ ch, float3 O, float3 D, SelfSlot self,
                               ChunkBest& best) {
    const GeomChunkHeader* h = reinterpret_cast<const GeomChunkHeader*>(ch);
    const int n_sphere = h->n_sphere, n_plane = h->n_plane, n_cuboid = h->n_cuboid, n_tri = h->n_tri;
    {
        const float4* sp = ch + h->off_sphere;
        int i = 0;
        if (n_sphere >= 16) {
            const int bulk = n_sphere & ~3;
#pragma unroll 4
            for (; i < bulk; ++i) sp_item_sphere(sp[i], O, D, i == self.sphere, self.mode, i, best);
        }
#pragma unroll 1
        for (; i < n_sphere; ++i) sp_item_sphere(sp[i], O, D, i == self.sphere, self.mode, i, best);
    }
    {
        const float4* pl = ch + h->off_plane;
#pragma unroll 1
        for (int i = 0; i < n_plane; ++i)
            sp_item_plane(pl[4 * i], pl[4 * i + 1], pl[4 * i + 2], pl[4 * i + 3], O, D, i == self.plane, n_sphere + i, best);
    }
    {
        const float4* cb = ch + h->off_cuboid;
#pragma unroll 1
        for (int i = 0; i < n_cuboid; ++i)
            sp_item_cuboid(cb[5 * i], cb[5 * i + 1], cb[5 * i + 2], cb[5 * i + 3], cb[5 * i + 4], O, D, i == self.cuboid,
                           self.mode, n_sphere + n_plane + i, best);
    }
    {
        const float4* tr = ch + h->off_tri;
        int i = 0;
        if (n_tri >= 16) {
            const int bulk = n_tri & ~1;
#pragma unroll 2
            for (; i < bulk; ++i)
                sp_item_triangle(tr[3 * i], tr[3 * i + 1], tr[3 * i + 2], O, D, i == self.tri, n_sphere + n_plane + n_cuboid + i, best);
        }
#pragma unroll 1
        for (; i < n_tri; ++i)
            sp_item_triangle(tr[3 * i], tr[3 * i + 1], tr[3 * i + 2], O, D, i == self.tri, n_sphere + n_plane + n_cuboid + i, best);
    }
    {
        const int n_aax = h->n_aax, n_aay = h->n_aay, n_aaz = h->n_aaz;
        if (n_aax + n_aay + n_aaz > 0) {
            const float4* aa = ch + h->off_aa;
            const int id_base = n_sphere + n_plane + n_cuboid + n_tri;
            if (n_aax > 0) sp_intersect_aa<0>(aa, 0, n_aax, id_base, O, D, fast_rcp(D.x), self.aa, best);
            if (n_aay > 0) sp_intersect_aa<1>(aa, n_aax, n_aay, id_base, O, D, fast_rcp(D.y), self.aa, best);
            if (n_aaz > 0) sp_intersect_aa<2>(aa, n_aax + n_aay, n_aaz, id_base, O, D, fast_rcp(D.z), self.aa, best);
        }
    }
}

// ---- lean walk for the warp-autonomous kernel (sp_warp_kernel.cuh) ---------------------------------------
// The same tests in the same order, with less bookkeeping per test: the ray's source collider is named by its
// position in the chunk's id array ("tag"), the winner is kept as one word  tag | (inner face) << 31, and an
// axis-aligned rectangle brings that word along (packed by the host), so its orientation costs one XOR with
// the sign of the ray's direction component.
SP_DEV void sp_lean_sphere(float4 s, float3 O, float3 D, bool is_self, uint32_t mode, uint32_t tag, float& bt, uint32_t& bcode) {
    float3 oc = O - xyz(s);
    float b = dot(D, oc);
    float3 q = fma3(D, -b, oc);
    float disc = s.w - dot(q, q);
    if (disc > 0.f) {
        float sq = fast_sqrt(disc);
        float h0 = -b - sq, h1 = -b + sq;
        bool near_ok = (h0 > 0.f) && !is_self;             // SP_SELF_FAR: only the far root
        float t = near_ok ? h0 : h1;
        bool ok = (t > 0.f) && !(is_self && mode != SP_SELF_FAR) && (t < bt);
        if (ok) { bt = t; bcode = near_ok ? tag : (tag | 0x80000000u); }
    }
}

SP_DEV void sp_lean_plane(float4 a, float4 c, float4 u4, float4 v4, float3 O, float3 D, bool is_self, uint32_t tag,
                          float& bt, uint32_t& bcode) {
    float3 N = xyz(a), oc = O - xyz(c);
    float nd = dot(N, D);
    nd = (nd == 0.f) ? 1e-4f : nd;
    float k = -dot(N, oc);
    float t = __fdividef(k, nd);
    float u = fmaf(t, dot(xyz(u4), D), dot(xyz(u4), oc));
    float v = fmaf(t, dot(xyz(v4), D), dot(xyz(v4), oc));
    bool ok = (fabsf(u) <= a.w) && (fabsf(v) <= c.w) && (k * nd > 0.f) && !is_self && (t < bt);
    if (ok) { bt = t; bcode = nd < 0.f ? tag : (tag | 0x80000000u); }
}

// The winner's code also names the slab the ray crosses at the hit, bits [8:10): the face normal is then a column of
// the inverse basis, oriented against the ray (sp_warp_kernel.cuh writes it into the fan record; locating the face from
// the hit point, as sp_collider_normal does, cost 45 instructions with a sixth of the lanes).
SP_DEV void sp_lean_cuboid(float4 r0, float4 r1, float4 r2, float4 c, float4 e, float3 O, float3 D, bool is_self,
                           uint32_t mode, uint32_t tag, float& bt, uint32_t& bcode) {
    float3 oc = O - xyz(c);
    float3 Ol = v3(dot(xyz(r0), oc), dot(xyz(r1), oc), dot(xyz(r2), oc));
    float3 Dl = v3(dot(xyz(r0), D), dot(xyz(r1), D), dot(xyz(r2), D));
    float ix = fast_rcp(Dl.x), iy = fast_rcp(Dl.y), iz = fast_rcp(Dl.z);   // +-inf for axis-parallel rays, as 1/0
    float t1 = (r0.w - Ol.x) * ix, t2 = (c.w - Ol.x) * ix;
    float t3 = (r1.w - Ol.y) * iy, t4 = (e.x - Ol.y) * iy;
    float t5 = (r2.w - Ol.z) * iz, t6 = (e.y - Ol.z) * iz;
    const float nx = fminf(t1, t2), ny = fminf(t3, t4), nz = fminf(t5, t6);
    const float fx = fmaxf(t1, t2), fy = fmaxf(t3, t4), fz = fmaxf(t5, t6);
    float tmin = fmaxf(fmaxf(nx, ny), nz);
    float tmax = fminf(fminf(fx, fy), fz);
    bool miss = (tmax < 0.f) || (tmin > tmax);
    bool inside = (tmin < 0.f) || is_self;                 // SP_SELF_FAR: exit point only
    float t = inside ? tmax : tmin;
    bool ok = !miss && !(is_self && mode != SP_SELF_FAR) && (t < bt);
    if (ok) {
        const uint32_t axis = inside ? (tmax == fx ? 0u : (tmax == fy ? 1u : 2u)) : (tmin == nx ? 0u : (tmin == ny ? 1u : 2u));
        bt = t; bcode = (inside ? (tag | 0x80000000u) : tag) | (axis << 8);
    }
}

SP_DEV void sp_lean_triangle(float4 m0, float4 m1, float4 m2, float3 O, float3 D, bool is_self, uint32_t tag, float& bt,
                             uint32_t& bcode) {
    float nd = dot(xyz(m2), D);
    nd = (nd == 0.f) ? 1e-4f : nd;
    const float w0 = dot(xyz(m2), O) + m2.w;
    const float t = -w0 * fast_rcp(nd);
    const float u = fmaf(t, dot(xyz(m0), D), dot(xyz(m0), O) + m0.w);
    const float v = fmaf(t, dot(xyz(m1), D), dot(xyz(m1), O) + m1.w);
    const bool ok = (u >= 0.f) && (v >= 0.f) && (u + v <= 1.f) && (t > 0.f) && !is_self && (t < bt);
    if (ok) { bt = t; bcode = nd < 0.f ? tag : (tag | 0x80000000u); }
}

#ifndef SP_LEAN_AA_UNROLL
#define SP_LEAN_AA_UNROLL 1
#endif
constexpr int kLeanAaUnroll = SP_LEAN_AA_UNROLL;
template <int A>
SP_DEV void sp_lean_aa(const float4* __restrict__ p, const float4* __restrict__ end, const float4* __restrict__ self_p,
                       float3 O, float3 D, float& bt, uint32_t& bcode) {
    const float oa = A == 0 ? O.x : (A == 1 ? O.y : O.z), da = A == 0 ? D.x : (A == 1 ? D.y : D.z);
    const float ob = A == 0 ? O.y : O.x, db = A == 0 ? D.y : D.x;
    const float oc = A == 2 ? O.y : O.z, dc = A == 2 ? D.y : D.z;
    const float inv_da = fast_rcp(da);                        // 1/0 = inf sends the hit point out of bounds
    const uint32_t flip = __float_as_uint(da) & 0x80000000u;
#pragma unroll kLeanAaUnroll
    for (; p != end; p += 2) {
        const float4 r0 = p[0], r1 = p[1];
        const float ca = A == 0 ? r0.x : (A == 1 ? r0.y : r0.z);
        const float cb = A == 0 ? r0.y : r0.x, cc = A == 2 ? r0.y : r0.z;
        const float t = (ca - oa) * inv_da;
        const float pb = fmaf(t, db, ob) - cb, pc = fmaf(t, dc, oc) - cc;
        const bool ok = (fabsf(pb) <= r1.x) && (fabsf(pc) <= r1.y) && (t > 0.f) && (t < bt) && (p != self_p);
        if (ok) { bt = t; bcode = __float_as_uint(r1.z) ^ flip; }
    }
}

// Nearest hit over one staged chunk: bt = distance (+inf: none), bcode = tag | (inner face) << 31.
SP_DEV void sp_intersect_lean(const float4* __restrict__ ch, float3 O, float3 D, int self_tag, uint32_t mode, float& bt,
                              uint32_t& bcode) {
    const int4 hn = *reinterpret_cast<const int4*>(ch);            // n_sphere n_plane n_cuboid n_tri
    const int4 ho = *reinterpret_cast<const int4*>(ch + 1);        // off_sphere off_plane off_cuboid off_tri
    int tag = 0;
    {
        const float4* sp = ch + ho.x;
#pragma unroll 1
        for (int i = 0; i < hn.x; ++i) sp_lean_sphere(sp[i], O, D, i == self_tag, mode, (uint32_t)i, bt, bcode);
        tag += hn.x;
    }
    if (hn.y > 0) {
        const float4* pl = ch + ho.y;
#pragma unroll 1
        for (int i = 0; i < hn.y; ++i)
            sp_lean_plane(pl[4 * i], pl[4 * i + 1], pl[4 * i + 2], pl[4 * i + 3], O, D, tag + i == self_tag, (uint32_t)(tag + i), bt, bcode);
        tag += hn.y;
    }
    {
        const float4* cb = ch + ho.z;
#pragma unroll 1
        for (int i = 0; i < hn.z; ++i)
            sp_lean_cuboid(cb[5 * i], cb[5 * i + 1], cb[5 * i + 2], cb[5 * i + 3], cb[5 * i + 4], O, D, tag + i == self_tag,
                           mode, (uint32_t)(tag + i), bt, bcode);
        tag += hn.z;
    }
    if (hn.w > 0) {
        const float4* tr = ch + ho.w;
#pragma unroll 1
        for (int i = 0; i < hn.w; ++i)
            sp_lean_triangle(tr[3 * i], tr[3 * i + 1], tr[3 * i + 2], O, D, tag + i == self_tag, (uint32_t)(tag + i), bt, bcode);
        tag += hn.w;
    }
    {
        const int4 ha = *reinterpret_cast<const int4*>(ch + 2);    // off_ids n_vec4 n_aax n_aay
        const int4 hb = *reinterpret_cast<const int4*>(ch + 3);    // n_aaz off_aa - -
        const float4* aa = ch + hb.y;
        const float4* self_p = aa + 2 * (self_tag - tag);         // outside the section if the source is not a rectangle
        const float4* e0 = aa + 2 * ha.z;
        const float4* e1 = e0 + 2 * ha.w;
        const float4* e2 = e1 + 2 * hb.x;
        sp_lean_aa<0>(aa, e0, self_p, O, D, bt, bcode);
        sp_lean_aa<1>(e0, e1, self_p, O, D, bt, bcode);
        sp_lean_aa<2>(e1, e2, self_p, O, D, bt, bcode);
    }
}

// ---- bounding-volume hierarchy over the small colliders of a large scene ---------------------------------
// An acceleration structure (the reference itself notes that meshes need one,
// triangle_mesh.py:7-9), used when a scene has SP_BVH_MIN_COLLIDERS colliders or more: with it a ray no longer
// tests every collider, so the bound of such scenes is memory latency / divergence, not FMA throughput.
// The colliders a ray does reach go through the very same sp_item_* tests, and the boxes are conservative,
// so hits are those of the exhaustive loop.  Colliders that span a large part of the scene (ground planes,
// sky boxes) stay in the staged chunk and are tested by every ray.
//   node  (4 float4): child 0 box (lo.xyz, hi.x | hi.yz) child 1 box (lo.xy | lo.z, hi.xyz), children (int, int)
//                     child >= 0: node index; child < 0: leaf (see below)
//   leaf  : ~child = float4 offset of its first record in `data` << 3 | (count - 1); a record is one header vector
//           (stream type | casts shadow << 8, collider id, number of data vectors, -) followed by the packed collider

// Slab test with the ray's origin folded into the multiply-add:  (lo - O) / D  =  lo * inv + noi  with inv = 1 / D and
// noi = -O * inv, six FFMA per box instead of six FADD + six FMUL.  inv comes from sp_safe_rcp (a zero direction component
// becomes +-1e-30, so inv stays finite and the products never meet inf - inf); the rounding of the folded form, 2^-24 of
// |lo * inv|, is far inside the boxes' relative padding of 1e-5 (sp_api.cu, collider_aabb).
SP_DEV float sp_safe_rcp(float d) { return fast_rcp(fabsf(d) > 1e-30f ? d : copysignf(1e-30f, d)); }
SP_DEV bool sp_box_hit(float3 lo, float3 hi, float3 inv, float3 noi, float t_max, float& t_near) {
    const float tx1 = fmaf(lo.x, inv.x, noi.x), tx2 = fmaf(hi.x, inv.x, noi.x);
    const float ty1 = fmaf(lo.y, inv.y, noi.y), ty2 = fmaf(hi.y, inv.y, noi.y);
    const float tz1 = fmaf(lo.z, inv.z, noi.z), tz2 = fmaf(hi.z, inv.z, noi.z);
    t_near = fmaxf(fmaxf(fminf(tx1, tx2), fminf(ty1, ty2)), fmaxf(fminf(tz1, tz2), 0.f));
    const float t_far = fminf(fminf(fmaxf(tx1, tx2), fmaxf(ty1, ty2)), fminf(fmaxf(tz1, tz2), t_max));
    return t_near <= t_far;
}

// Nearest hit among the BVH's colliders closer than best.t.  casters_only: shadow rays (glossy.py:53-57) look at
// shadow-casting colliders only and may stop at the first hit closer than t_any.
//
// "while-while" traversal (Aila & Laine, Understanding the efficiency of ray traversal on GPUs): an inner loop walks
// box nodes until the lane holds a leaf, a second loop tests the leaf's colliders.  Lanes that reach their leaf early
// wait at the end of the inner loop, so the leaves of a warp are tested side by side instead of one lane at a time
// between the node steps of the others (round 1's single loop ran its collider tests with 2 of 32 lanes, ncu
// profiles/r2_stress_bvh_binary.md).  The host sorts the colliders of a leaf by type, so the lanes of a warp mostly
// agree on the test they run for item k of their leaves.
#define SP_BVH_DONE 0x7FFFFFFF
SP_DEV void sp_bvh_nearest(const DBvh& bvh, float3 O, float3 D, int src_id, uint32_t mode, bool casters_only, float t_any,
                           ChunkBest& best) {
    if (bvh.n_nodes == 0) return;
    const float3 inv = v3(sp_safe_rcp(D.x), sp_safe_rcp(D.y), sp_safe_rcp(D.z));
    const float3 noi = v3(-O.x * inv.x, -O.y * inv.y, -O.z * inv.z);
    int stack[32];
    int sp = 0, node = 0;                                  // >= 0: box node, < 0: leaf code, SP_BVH_DONE: finished
    while (node != SP_BVH_DONE) {
        // ---- box nodes until a leaf (or nothing) is at hand ----------------------------------------------------
        while (node >= 0 && node != SP_BVH_DONE) {
            const float4 n0 = __ldg(bvh.nodes + 4 * node), n1 = __ldg(bvh.nodes + 4 * node + 1);
            const float4 n2 = __ldg(bvh.nodes + 4 * node + 2), n3 = __ldg(bvh.nodes + 4 * node + 3);
            float ta, tb;
            const bool ha = sp_box_hit(v3(n0.x, n0.y, n0.z), v3(n0.w, n1.x, n1.y), inv, noi, best.t, ta);
            const bool hb = sp_box_hit(v3(n1.z, n1.w, n2.x), v3(n2.y, n2.z, n2.w), inv, noi, best.t, tb);
            int ca = __float_as_int(n3.x), cb = __float_as_int(n3.y);
            if (ha && hb) {
                if (tb < ta) { const int c = ca; ca = cb; cb = c; }      // nearer child first
                stack[sp++] = cb;
                node = ca;
            } else if (ha || hb) {
                node = ha ? ca : cb;
            } else {
                node = sp ? stack[--sp] : SP_BVH_DONE;
            }
        }
        // ---- the leaf's colliders ------------------------------------------------------------------------------
        if (node != SP_BVH_DONE) {
            const int code = ~node, count = (code & 7) + 1;
            const float4* rec = bvh.data + (code >> 3);        // leaf records: header vector + packed collider, back to back
            for (int i = 0; i < count; ++i) {
                const float4 head = __ldg(rec);
                const int kind = __float_as_int(head.x), id = __float_as_int(head.y);
                const float4* d = rec + 1;
                rec = d + __float_as_int(head.z);
                if (casters_only && !(kind & 256)) continue;
                const bool is_self = id == src_id;
                switch (kind & 255) {
                case SP_ST_SPHERE: sp_item_sphere(__ldg(d), O, D, is_self, mode, id, best); break;
                case SP_ST_PLANE: sp_item_plane(__ldg(d), __ldg(d + 1), __ldg(d + 2), __ldg(d + 3), O, D, is_self, id, best); break;
                case SP_ST_CUBOID: sp_item_cuboid(__ldg(d), __ldg(d + 1), __ldg(d + 2), __ldg(d + 3), __ldg(d + 4), O, D, is_self, mode, id, best); break;
                case SP_ST_TRI: sp_item_triangle(__ldg(d), __ldg(d + 1), __ldg(d + 2), O, D, is_self, id, best); break;
                case SP_ST_AAX: sp_item_aa<0>(__ldg(d), __ldg(d + 1), O, D, inv.x, is_self, id, best); break;
                case SP_ST_AAY: sp_item_aa<1>(__ldg(d), __ldg(d + 1), O, D, inv.y, is_self, id, best); break;
                default: sp_item_aa<2>(__ldg(d), __ldg(d + 1), O, D, inv.z, is_self, id, best); break;
                }
            }
            if (best.t < t_any) return;
            node = sp ? stack[--sp] : SP_BVH_DONE;
        }
    }
}

SP_DEV int sp_chunk_id(const float4* __restrict__ ch, int idx) {
    const GeomChunkHeader* h = reinterpret_cast<const GeomChunkHeader*>(ch);
    return reinterpret_cast<const int*>(ch + h->off_ids)[idx];
}
