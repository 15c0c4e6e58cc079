// Primary-ray generation and importance-sampled diffuse directions.
//
// Restates Camera.get_ray (camera.py:51-85), cosine_pdf / spherical_caps_pdf / mixed_pdf
// (random.py:50-174).  The reference draws 2 + 3 uniforms per direction and evaluates both pdfs'
// generators for every ray before selecting one; here one Philox block gives
//   u0: mixture choice   u1: azimuth   u2: radial variate   u3: which cap
// and only the selected generator runs (same distribution; oracle/ mirrors this mapping).
#pragma once
#include "sp_rng.cuh"
#include "sp_types.cuh"

// tangent frame of random.py:60-63: a = (0,1,0) if |w.x| > 0.9 else (1,0,0); v = norm(w x a); u = w x v
SP_DEV void sp_onb(float3 w, float3& u, float3& v) {
    float3 a = fabsf(w.x) > 0.9f ? v3(0.f, 1.f, 0.f) : v3(1.f, 0.f, 0.f);
    v = normalize0(cross(w, a));
    u = cross(w, v);
}

SP_DEV void sp_camera_ray(const DCamera& cam, uint32_t pixel, uint32_t sample, const uint32_t* __restrict__ keys,
                          float3& origin, float3& dir) {
    float u[4];
    sp_draw4_keys(pixel, sp_root_path(sample), SP_BLOCK_DIRECTION, keys, u);
    const uint32_t py = cam.W > 1 ? (uint32_t)__umul64hi((unsigned long long)pixel, cam.w_magic) : pixel;
    const uint32_t px = pixel - py * (uint32_t)cam.W;
    // np.linspace(-w/2, w/2, W)[px] and np.linspace(h/2, -h/2, H)[py]
    const float gx = fmaf((float)px, cam.step_x, -0.5f * cam.cam_w);
    const float gy = fmaf(-(float)py, cam.step_y, 0.5f * cam.cam_h);
    const float x = fmaf(u[0] - 0.5f, cam.jitter_x, gx);
    const float y = fmaf(u[1] - 0.5f, cam.jitter_y, gy);
    origin = cam.look_from;
    if (cam.lens_radius != 0.f) {                          // thin lens: a point of the aperture disk (random_in_unit_disk)
        float rr = sqrtf(u[2]), sn, cs;
        sincospif(2.f * u[3], &sn, &cs);
        const float rx = rr * cs * cam.lens_radius, ry = rr * sn * cam.lens_radius;
        origin = cam.look_from + cam.right * rx + cam.up * ry;
    }
    const float fd = cam.focal_distance;
    const float3 target = cam.look_from + cam.up * (y * fd) + cam.right * (x * fd) + cam.fwd * fd;
    dir = normalize0(target - origin);
}

// Sample a direction for a diffuse bounce at `origin` with shading normal N; returns the estimator
// weight  clip(N.d, 0, 1) / pdf(d) / pi  (diffuse.py:76-81), 0 if the sample carries nothing.
// `imp`: the importance list (centre.xyz, radius).  sc.importance lives in the kernel-parameter bank, where the
// per-lane index of the picked cap serialises the access; callers that keep a copy in shared memory pass that.
template <typename ImpList>
SP_DEV float sp_sample_diffuse_with(const DScene& sc, const ImpList& imp, float3 origin, float3 N, float w_cos, uint32_t pix,
                                    uint32_t path, float3& dir) {
    float u[4];
    sp_draw4_keys(pix, path, SP_BLOCK_DIRECTION, sc.philox_keys, u);
    float sn, cs;
    fast_sincos_2pi(u[1], sn, cs);
    const int l = sc.n_importance;
    bool use_cos = (l == 0) || (u[0] < w_cos);
    // both generators build the direction the same way (random.py:60-71 / 128-148): polar angle about an
    // axis w, azimuth phi in the frame (u, v) of that axis; only the axis and the polar law differ, so the
    // frame and the direction are computed once for the whole warp
    float3 w = N;
    float z, s;
    if (use_cos) {                                           // cosine_pdf.generate
        s = fast_sqrt(u[2]);
        z = fast_sqrt(1.f - u[2]);
    } else {                                                 // spherical_caps_pdf.generate
        int pick = min((int)(u[3] * (float)l), l - 1);
        const float4 ip = imp(pick);
        float3 to_c = xyz(ip) - origin;
        float d2 = dot(to_c, to_c);
        float inv = rsqrtf(d2);
        w = to_c * inv;
        float ratio = clamp01(ip.w * inv);
        float cmax = fast_sqrt(1.f - ratio * ratio);
        z = 1.f + u[2] * (cmax - 1.f);
        s = fast_sqrt(fmaxf(1.f - z * z, 0.f));
    }
    float3 au, av;
    sp_onb(w, au, av);
    dir = au * (cs * s) + av * (sn * s) + w * z;
    float ndl = clamp01(dot(dir, N));
    if (ndl <= 0.f) return 0.f;
    float pdf = ndl * (1.f / SP_PI);
    if (l > 0) {                                             // mixed_pdf.value
        float caps = 0.f;
#pragma unroll 2
        for (int i = 0; i < l; ++i) {
            const float4 ii = imp(i);
            float3 to_c = xyz(ii) - origin;
            float d2 = dot(to_c, to_c);
            float inv = rsqrtf(d2);
            float ratio = clamp01(ii.w * inv);
            float cmax = fast_sqrt(1.f - ratio * ratio);
            if (dot(dir, to_c) * inv > cmax) caps += __fdividef(1.f, (1.f - cmax) * 2.f * SP_PI);
        }
        pdf = pdf * w_cos + (caps * sc.inv_n_importance) * (1.f - w_cos);
    }
    return __fdividef(ndl, pdf) * (1.f / SP_PI);
}

SP_DEV float sp_sample_diffuse(const DScene& sc, float3 origin, float3 N, float w_cos, uint32_t pix,
                               uint32_t path, float3& dir) {
    auto imp = [&](int i) { return make_float4(sc.importance[i].center.x, sc.importance[i].center.y, sc.importance[i].center.z,
                                               sc.importance[i].radius); };
    return sp_sample_diffuse_with(sc, imp, origin, N, w_cos, pix, path, dir);
}
