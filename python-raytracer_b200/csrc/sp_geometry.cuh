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

// One axis-aligned rectangle section: normal along axis A (0, 1, 2), in-plane axes B < C.
template <int A>
SP_DEV void sp_intersect_aa(const float4* __restrict__ aa, int first, int count, int id_base, float3 O, float3 D,
                            float inv_da, int self_aa, ChunkBest& best) {
    const float oa = A == 0 ? O.x : (A == 1 ? O.y : O.z), da = A == 0 ? D.x : (A == 1 ? D.y : D.z);
    const float ob = A == 0 ? O.y : O.x, db = A == 0 ? D.y : D.x;
    const float oc = A == 2 ? O.y : O.z, dc = A == 2 ? D.y : D.z;
#pragma unroll 2
    for (int i = first; i < first + count; ++i) {
        const float4 r0 = aa[2 * i], r1 = aa[2 * i + 1];
        const float ca = A == 0 ? r0.x : (A == 1 ? r0.y : r0.z);
        const float cb = A == 0 ? r0.y : r0.x, cc = A == 2 ? r0.y : r0.z;
        const float t = (ca - oa) * inv_da;                        // k / N.D with the normal's sign cancelled
        const float pb = fmaf(t, db, ob) - cb, pc = fmaf(t, dc, oc) - cc;
        const bool ok = (fabsf(pb) <= r1.x) && (fabsf(pc) <= r1.y) && (t > 0.f) && (i != self_aa);
        if (ok && t < best.t) { best.t = t; best.idx = id_base + i; best.orient = (r0.w * da < 0.f) ? 1 : -1; }
    }
}

SP_DEV void sp_intersect_chunk(const float4* __restrict__ ch, float3 O, float3 D, SelfSlot self,
                               ChunkBest& best) {
    const GeomChunkHeader* h = reinterpret_cast<const GeomChunkHeader*>(ch);
    const int n_sphere = h->n_sphere, n_plane = h->n_plane, n_cuboid = h->n_cuboid, n_tri = h->n_tri;

    // ---- spheres ---------------------------------------------------------------------------
    {
        const float4* sp = ch + h->off_sphere;
#pragma unroll 4
        for (int i = 0; i < n_sphere; ++i) {
            float4 s = sp[i];
            float3 oc = O - xyz(s);
            float b = dot(D, oc);
            float3 q = fma3(D, -b, oc);
            float disc = s.w - dot(q, q);
            if (disc > 0.f) {
                float sq = fast_sqrt(disc);
                float h0 = -b - sq, h1 = -b + sq;
                bool is_self = (i == self.sphere);
                bool near_ok = (h0 > 0.f) && !is_self;             // SP_SELF_FAR: only the far root
                float t = near_ok ? h0 : h1;
                bool ok = (t > 0.f) && !(is_self && self.mode != SP_SELF_FAR);
                if (ok && t < best.t) { best.t = t; best.idx = i; best.orient = near_ok ? 1 : -1; }
            }
        }
    }
    // ---- bounded planes ------------------------------------------------------------------------
    {
        const float4* pl = ch + h->off_plane;
#pragma unroll 2
        for (int i = 0; i < n_plane; ++i) {
            float4 a = pl[4 * i], c = pl[4 * i + 1], u4 = pl[4 * i + 2], v4 = pl[4 * i + 3];
            float3 N = xyz(a), oc = O - xyz(c);
            float nd = dot(N, D);
            nd = (nd == 0.f) ? 1e-4f : nd;
            float k = -dot(N, oc);
            float t = __fdividef(k, nd);
            float u = fmaf(t, dot(xyz(u4), D), dot(xyz(u4), oc));
            float v = fmaf(t, dot(xyz(v4), D), dot(xyz(v4), oc));
            bool ok = (fabsf(u) <= a.w) && (fabsf(v) <= c.w) && (k * nd > 0.f) && (i != self.plane);
            if (ok && t < best.t) { best.t = t; best.idx = n_sphere + i; best.orient = nd < 0.f ? 1 : -1; }
        }
    }
    // ---- oriented cuboids ------------------------------------------------------------------------
    {
        const float4* cb = ch + h->off_cuboid;
        for (int i = 0; i < n_cuboid; ++i) {
            float4 r0 = cb[5 * i], r1 = cb[5 * i + 1], r2 = cb[5 * i + 2], c = cb[5 * i + 3], e = cb[5 * i + 4];
            float3 oc = O - xyz(c);
            float3 Ol = v3(dot(xyz(r0), oc), dot(xyz(r1), oc), dot(xyz(r2), oc));
            float3 Dl = v3(dot(xyz(r0), D), dot(xyz(r1), D), dot(xyz(r2), D));
            float ix = fast_rcp(Dl.x), iy = fast_rcp(Dl.y), iz = fast_rcp(Dl.z);   // +-inf for axis-parallel rays, as 1/0
            float t1 = (r0.w - Ol.x) * ix, t2 = (c.w - Ol.x) * ix;
            float t3 = (r1.w - Ol.y) * iy, t4 = (e.x - Ol.y) * iy;
            float t5 = (r2.w - Ol.z) * iz, t6 = (e.y - Ol.z) * iz;
            float tmin = fmaxf(fmaxf(fminf(t1, t2), fminf(t3, t4)), fminf(t5, t6));
            float tmax = fminf(fminf(fmaxf(t1, t2), fmaxf(t3, t4)), fmaxf(t5, t6));
            bool is_self = (i == self.cuboid);
            bool miss = (tmax < 0.f) || (tmin > tmax);
            bool inside = (tmin < 0.f) || is_self;                 // SP_SELF_FAR: exit point only
            float t = inside ? tmax : tmin;
            bool ok = !miss && !(is_self && self.mode != SP_SELF_FAR);
            if (ok && t < best.t) { best.t = t; best.idx = n_sphere + n_plane + i; best.orient = inside ? -1 : 1; }
        }
    }
    // ---- triangles -----------------------------------------------------------------------------
    // triangle.py:37-66 through the affine map to the unit triangle: the three edge tests
    // n31.(M-p1) >= 0, n12.(M-p2) >= 0, n23.(M-p3) >= 0 are u >= 0, v >= 0, 1-u-v >= 0 of the hit point's
    // barycentric coordinates, and the third row of the map is the plane normal (N.D, N.(O-p1)).
    {
        const float4* tr = ch + h->off_tri;
#pragma unroll 2
        for (int i = 0; i < n_tri; ++i) {
            const float4 m0 = tr[3 * i], m1 = tr[3 * i + 1], m2 = tr[3 * i + 2];
            float nd = dot(xyz(m2), D);
            nd = (nd == 0.f) ? 1e-4f : nd;
            const float w0 = dot(xyz(m2), O) + m2.w;                 // N.(O - p1) = -k
            const float t = -w0 * fast_rcp(nd);
            const float u = fmaf(t, dot(xyz(m0), D), dot(xyz(m0), O) + m0.w);
            const float v = fmaf(t, dot(xyz(m1), D), dot(xyz(m1), O) + m1.w);
            const bool ok = (u >= 0.f) && (v >= 0.f) && (u + v <= 1.f) && (t > 0.f) && (i != self.tri);
            if (ok && t < best.t) {
                best.t = t; best.idx = n_sphere + n_plane + n_cuboid + i; best.orient = nd < 0.f ? 1 : -1;
            }
        }
    }
    // ---- axis-aligned rectangles -----------------------------------------------------------------
    {
        const int n_aax = h->n_aax, n_aay = h->n_aay, n_aaz = h->n_aaz;
        if (n_aax + n_aay + n_aaz > 0) {
            const float4* aa = ch + h->off_aa;
            const int id_base = n_sphere + n_plane + n_cuboid + n_tri;
            // 1/0 = inf sends the hit point of an axis-parallel ray out of bounds (the reference
            // substitutes N.D = 1e-4 there and misses all the same)
            if (n_aax > 0) sp_intersect_aa<0>(aa, 0, n_aax, id_base, O, D, fast_rcp(D.x), self.aa, best);
            if (n_aay > 0) sp_intersect_aa<1>(aa, n_aax, n_aay, id_base, O, D, fast_rcp(D.y), self.aa, best);
            if (n_aaz > 0) sp_intersect_aa<2>(aa, n_aax + n_aay, n_aaz, id_base, O, D, fast_rcp(D.z), self.aa, best);
        }
    }
}

SP_DEV int sp_chunk_id(const float4* __restrict__ ch, int idx) {
    const GeomChunkHeader* h = reinterpret_cast<const GeomChunkHeader*>(ch);
    return reinterpret_cast<const int*>(ch + h->off_ids)[idx];
}
