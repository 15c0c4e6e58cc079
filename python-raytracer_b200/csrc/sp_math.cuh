// Small float3 / complex helpers used by the device code.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SP_DEV __device__ __forceinline__
#define SP_PI 3.14159265358979323846f
#define SP_INF __int_as_float(0x7f800000)

SP_DEV float3 v3(float x, float y, float z) { return make_float3(x, y, z); }
SP_DEV float3 v3(float s) { return make_float3(s, s, s); }
SP_DEV float3 operator+(float3 a, float3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
SP_DEV float3 operator-(float3 a, float3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
SP_DEV float3 operator-(float3 a) { return v3(-a.x, -a.y, -a.z); }
SP_DEV float3 operator*(float3 a, float3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
SP_DEV float3 operator*(float3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
SP_DEV float3 operator*(float s, float3 a) { return v3(a.x * s, a.y * s, a.z * s); }
SP_DEV float3 operator/(float3 a, float3 b) { return v3(a.x / b.x, a.y / b.y, a.z / b.z); }
SP_DEV float3& operator+=(float3& a, float3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
SP_DEV float dot(float3 a, float3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
SP_DEV float3 cross(float3 a, float3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
SP_DEV float3 fma3(float3 a, float s, float3 b) {   // a*s + b
    return v3(fmaf(a.x, s, b.x), fmaf(a.y, s, b.y), fmaf(a.z, s, b.z));
}
// vec3.normalize (vector3.py:158-160): zero vectors stay zero
SP_DEV float3 normalize0(float3 a) {
    float m2 = dot(a, a);
    float inv = m2 > 0.f ? rsqrtf(m2) : 1.f;
    return a * inv;
}
// sqrt.approx / sin.approx / cos.approx (MUFU, ~1e-6 relative): used where the result feeds a random
// direction or a hit distance, never a texel index
SP_DEV float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }   // 1/0 = inf
SP_DEV float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// sin and cos of 2*pi*u for u in [0, 1): evaluated at 2*pi*(u - 0.5), where the MUFU approximations are
// accurate to ~5e-7 absolute, and negated (sin(x + pi) = -sin x)
SP_DEV void fast_sincos_2pi(float u, float& sn, float& cs) {
    const float x = (u - 0.5f) * 6.28318530717958647692f;
    sn = -__sinf(x); cs = -__cosf(x);
}
SP_DEV float clamp01(float x) { return fminf(fmaxf(x, 0.f), 1.f); }
SP_DEV float3 xyz(float4 a) { return v3(a.x, a.y, a.z); }
SP_DEV bool any_nonzero(float3 a) { return a.x != 0.f || a.y != 0.f || a.z != 0.f; }

// row-major 3x3 (9 floats) times vector
SP_DEV float3 mat3_mul(const float* m, float3 a) {
    return v3(fmaf(m[0], a.x, fmaf(m[1], a.y, m[2] * a.z)),
              fmaf(m[3], a.x, fmaf(m[4], a.y, m[5] * a.z)),
              fmaf(m[6], a.x, fmaf(m[7], a.y, m[8] * a.z)));
}

// ---- complex numbers (refractive indices are complex per colour channel) ----------------------
struct cplx { float re, im; };
SP_DEV cplx cx(float re, float im = 0.f) { cplx c; c.re = re; c.im = im; return c; }
SP_DEV cplx operator+(cplx a, cplx b) { return cx(a.re + b.re, a.im + b.im); }
SP_DEV cplx operator-(cplx a, cplx b) { return cx(a.re - b.re, a.im - b.im); }
SP_DEV cplx operator*(cplx a, cplx b) { return cx(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
SP_DEV cplx operator*(cplx a, float s) { return cx(a.re * s, a.im * s); }
SP_DEV cplx operator/(cplx a, cplx b) {
    float d = __frcp_rn(b.re * b.re + b.im * b.im);
    return cx((a.re * b.re + a.im * b.im) * d, (a.im * b.re - a.re * b.im) * d);
}
SP_DEV float cabs2(cplx a) { return a.re * a.re + a.im * a.im; }
// principal square root (numpy semantics incl. the sign of a zero imaginary part)
SP_DEV cplx csqrt(cplx z) {
    float m = sqrtf(cabs2(z));
    float re = sqrtf(fmaxf(0.5f * (m + z.re), 0.f));
    float im = sqrtf(fmaxf(0.5f * (m - z.re), 0.f));
    return cx(re, copysignf(im, z.im));
}
