// Geometry of a hit: exact distance, point, collider normal, uv and texel addressing.
//
// Templated on the arithmetic type.  The wavefront runs in float32; for materials whose colour
// depends on *which texel* a hit lands in (image textures, normal maps, the thin-film noise/LUT)
// the winning collider alone is re-intersected in float64 from the float32 ray and the double
// precision scene parameters (`precise` materials).  Nearest-neighbour, highly repeated textures
// amplify a 1e-7 relative error of the hit point into a different texel ("aliased
// textures"); a handful of double operations per *hit* — not per ray/collider test — removes
// that error source for primary hits.  B200 runs FP64 at half the FP32 rate, so this is cheap.
//
// Restates: sphere.py:54-64, plane.py:98-105, cuboid.py:142-187, triangle.py:85-86,
// texture.py:32-39 (negative-row indexing), cuboid.py:29-32 / skybox.py:29-32 (cross layout).
#pragma once
#include "sp_types.cuh"

template <typename T> struct tv3 { T x, y, z; };
template <typename T> SP_DEV tv3<T> mk(T x, T y, T z) { tv3<T> r; r.x = x; r.y = y; r.z = z; return r; }
template <typename T> SP_DEV tv3<T> operator+(tv3<T> a, tv3<T> b) { return mk<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> SP_DEV tv3<T> operator-(tv3<T> a, tv3<T> b) { return mk<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> SP_DEV tv3<T> operator*(tv3<T> a, T s) { return mk<T>(a.x * s, a.y * s, a.z * s); }
template <typename T> SP_DEV T tdot(tv3<T> a, tv3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> SP_DEV tv3<T> ld3(const T* p) { return mk<T>(p[0], p[1], p[2]); }
template <typename T> SP_DEV tv3<T> tmat(const T* m, tv3<T> a) {
    return mk<T>(m[0] * a.x + m[1] * a.y + m[2] * a.z, m[3] * a.x + m[4] * a.y + m[5] * a.z,
                 m[6] * a.x + m[7] * a.y + m[8] * a.z);
}
template <typename T> SP_DEV tv3<T> from_f3(float3 a) { return mk<T>((T)a.x, (T)a.y, (T)a.z); }
template <typename T> SP_DEV float3 to_f3(tv3<T> a) { return v3((float)a.x, (float)a.y, (float)a.z); }

// payload slots (include/sightpy_b200.h)
#define SP_SPH_C 0
#define SP_SPH_R 3
#define SP_PL_C 0
#define SP_PL_U 3
#define SP_PL_V 6
#define SP_PL_N 9
#define SP_PL_W 12
#define SP_PL_H 13
#define SP_PL_SHIFT 14
#define SP_PL_INVB 16
#define SP_CB_C 0
#define SP_CB_AW 3
#define SP_CB_AH 6
#define SP_CB_AL 9
#define SP_CB_LB 12
#define SP_CB_RT 15
#define SP_CB_SIZE 18
#define SP_CB_B 21
#define SP_CB_INVB 30
#define SP_TR_N 9
#define SP_TR_CEN 12
// derived reciprocals appended to the device copies of the payload by sp_scene_commit (divisions are
// the dearest instructions of the double-precision texel path)
#define SP_DEV_PAYLOAD 44
#define SP_SPH_INVR 4
#define SP_PL_INVW 25
#define SP_PL_INVH 26
#define SP_CB_INVSIZE 40

// Distance along the ray to collider `type`/`p`, choosing the root the float32 pass selected
// (orient: +1 near/outer, -1 far/inner).  Same formulas as the float32 tests, in T.
template <typename T>
SP_DEV T sp_refine_t(int type, const T* p, tv3<T> O, tv3<T> D, int orient, T t_fallback) {
    if (type == 0) {                                   // sphere (sphere.py:28-40, a = 1 assumed as upstream)
        tv3<T> oc = O - ld3(p + SP_SPH_C);
        T b = tdot(D, oc);
        T disc = b * b - (tdot(oc, oc) - p[SP_SPH_R] * p[SP_SPH_R]);
        if (!(disc > (T)0)) return t_fallback;
        T sq = sqrt(disc);
        return orient > 0 ? (-b - sq) : (-b + sq);
    } else if (type == 1 || type == 3) {               // plane / triangle: distance = |D * k / N.D|
        tv3<T> N = ld3(p + (type == 1 ? SP_PL_N : SP_TR_N));
        tv3<T> C = ld3(p + (type == 1 ? SP_PL_C : SP_TR_CEN));
        T nd = tdot(N, D);
        if (nd == (T)0) nd = (T)1e-4;
        return fabs(tdot(N, C - O) / nd) * sqrt(tdot(D, D));
    } else {                                           // cuboid
        const T* B = p + SP_CB_B;
        tv3<T> Ol = tmat(B, O), Dl = tmat(B, D);
        T ix = (T)1 / Dl.x, iy = (T)1 / Dl.y, iz = (T)1 / Dl.z;
        T t1 = (p[SP_CB_LB] - Ol.x) * ix, t2 = (p[SP_CB_RT] - Ol.x) * ix;
        T t3 = (p[SP_CB_LB + 1] - Ol.y) * iy, t4 = (p[SP_CB_RT + 1] - Ol.y) * iy;
        T t5 = (p[SP_CB_LB + 2] - Ol.z) * iz, t6 = (p[SP_CB_RT + 2] - Ol.z) * iz;
        T tmin = fmax(fmax(fmin(t1, t2), fmin(t3, t4)), fmin(t5, t6));
        T tmax = fmin(fmin(fmax(t1, t2), fmax(t3, t4)), fmax(t5, t6));
        return orient > 0 ? tmin : tmax;
    }
}

// Outward geometric normal of the collider at P (un-oriented).
template <typename T>
SP_DEV tv3<T> sp_collider_normal(int type, const T* p, tv3<T> P) {
    if (type == 0) return (P - ld3(p + SP_SPH_C)) * p[SP_SPH_INVR];
    if (type == 1) return ld3(p + SP_PL_N);
    if (type == 3) return ld3(p + SP_TR_N);
    tv3<T> Pl = tmat(p + SP_CB_B, P - ld3(p + SP_CB_C));
    T ax = fabs(Pl.x) * p[SP_CB_INVSIZE], ay = fabs(Pl.y) * p[SP_CB_INVSIZE + 1];
    T az = fabs(Pl.z) * p[SP_CB_INVSIZE + 2];
    T am = fmax(fmax(ax, ay), az);
    auto sgn = [](T v) { return v > (T)0 ? (T)1 : (v < (T)0 ? (T)-1 : (T)0); };
    tv3<T> face = mk<T>(am == ax ? sgn(Pl.x) : (T)0, am == ay ? sgn(Pl.y) : (T)0, am == az ? sgn(Pl.z) : (T)0);
    return tmat(p + SP_CB_INVB, face);
}

// Collider uv at P; `Nc` is the collider normal at P (needed by the cuboid's face select).
template <typename T>
SP_DEV void sp_collider_uv(int type, const T* p, tv3<T> P, tv3<T> Nc, bool cross_layout, T& u, T& v) {
    const T pi = (T)3.14159265358979323846;
    if (type == 0) {
        tv3<T> m = (P - ld3(p + SP_SPH_C)) * p[SP_SPH_INVR];
        u = (atan2(m.z, m.x) + pi) * (T)0.15915494309189533577;          // 1 / (2 pi)
        v = (asin(m.y) + pi / (T)2) * (T)0.31830988618379067154;           // 1 / pi
    } else if (type == 1) {
        tv3<T> mc = P - ld3(p + SP_PL_C);
        u = (tdot(ld3(p + SP_PL_U), mc) * p[SP_PL_INVW] + (T)1) * (T)0.5 + p[SP_PL_SHIFT];
        v = (tdot(ld3(p + SP_PL_V), mc) * p[SP_PL_INVH] + (T)1) * (T)0.5 + p[SP_PL_SHIFT + 1];
    } else if (type == 2) {
        tv3<T> mc = P - ld3(p + SP_CB_C);
        const T k = p[SP_CB_INVSIZE] * (T)(2 * 0.985);       // every face is scaled by the box width (cuboid.py:165-186)
        T dw = tdot(ld3(p + SP_CB_AW), mc), dh = tdot(ld3(p + SP_CB_AH), mc), dl = tdot(ld3(p + SP_CB_AL), mc);
        auto g = [k](T d, T off) { return (d * k + (T)1) * (T)0.5 + off; };
        auto is = [Nc](T x, T y, T z) { return Nc.x == x && Nc.y == y && Nc.z == z; };
        u = (T)0; v = (T)0;                              // rotated boxes match no face (cuboid.py:157-162)
        if (is(0, -1, 0))      { u = g(dw, 1);  v = g(-dl, 0); }   // BOTTOM
        else if (is(0, 1, 0))  { u = g(dw, 1);  v = g(dl, 2); }    // TOP
        else if (is(1, 0, 0))  { u = g(dl, 2);  v = g(dh, 1); }    // RIGHT
        else if (is(-1, 0, 0)) { u = g(-dl, 0); v = g(dh, 1); }    // LEFT
        else if (is(0, 0, 1))  { u = g(-dw, 3); v = g(dh, 1); }    // FRONT
        else if (is(0, 0, -1)) { u = g(dw, 1);  v = g(dh, 1); }    // BACK
    } else {
        u = (T)0; v = (T)0;                              // triangles have no uv mapping upstream
    }
    if (cross_layout) { u = u * (T)0.25; v = v * (T)(1.0 / 3.0); }
}

// a % n with Python's sign convention (result in [0, n)), n > 0.  Indices below 2^23 in magnitude — every texture
// that fits in memory — go through a float reciprocal: the quotient estimate is off by at most one, which the two
// fix-ups absorb (9 instructions instead of the ~22 of a 32-bit integer remainder by a run-time divisor).
SP_DEV int sp_pymod(int a, int n) {
    int r;
    if (abs(a) < (1 << 23)) {
        const int q = (int)((float)a * __frcp_rn((float)n));
        r = a - q * n;
        if (r < 0) r += n; else if (r >= n) r -= n;
        if (r < 0) r += n;
    } else {
        r = a % n; if (r < 0) r += n;
    }
    return r;
}

// img[-(int(v*H*repeat) % H), int(u*W*repeat) % W] with Python's floor-mod and negative indexing.
template <typename T>
SP_DEV int sp_texel_offset(T u, T v, int H, int W, T repeat, int Hreal, int Wreal) {
    const T fv = v * (T)H * repeat, fu = u * (T)W * repeat;
    long long r, c;
    if (fabs(fv) < (T)2.0e9 && fabs(fu) < (T)2.0e9) {   // the usual case: 32-bit remainders (64-bit ones cost ~100 instructions)
        const int iv = (int)fv, iu = (int)fu;           // astype(int): truncation towards zero
        r = sp_pymod(iv, H); c = sp_pymod(iu, W);       // Python % with positive modulus
    } else {
        const long long iv = (long long)fv, iu = (long long)fu;
        r = iv % H; if (r < 0) r += H;
        c = iu % W; if (c < 0) c += W;
    }
    long long row = (r == 0) ? 0 : (long long)Hreal - r;   // negative index -r of the indexed array
    row = row < 0 ? 0 : (row >= Hreal ? Hreal - 1 : row);  // (the reference would raise IndexError)
    c = c >= Wreal ? Wreal - 1 : c;
    return (int)(row * Wreal + c);
}
