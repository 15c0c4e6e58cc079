// sm_100a kernels of the sightpy backend: the fused wavefront level kernel (generate -> intersect
// -> shade -> emit), the frame resolve (average + sRGB tonemap) and two roofline micro-benchmarks.
//
// One *level launch* consumes every ray of one recursion depth of the reference's get_raycolor
// tree (ray.py:122-148) for the current chunk of primaries:
//   level 0      rays are generated in registers from (pixel, sample) — Camera.get_ray,
//                camera.py:51-85 — or read from caller arrays (sp_trace);
//   level >= 1   rays come from the previous level's queues: explicit ray records (reflection /
//                refraction / transmission children) and "fan" records (a diffuse hit: origin +
//                shading normal), each of which expands into diffuse_rays importance-sampled
//                directions *inside this kernel*, so the 20 children of diffuse.py:34-47 never
//                exist in memory.
// The launch is persistent: grid = SMs x resident CTAs, CTAs stride over the work items, the item
// count is read from device counters, so a whole chunk (all levels) is enqueued with no host
// round trip.  The colliders are staged into shared memory in 32 KB type-sorted chunks and every
// lane of a warp reads the same address (broadcast).
#include <cstdio>

#include "sp_launch.h"
#include "sp_sampling.cuh"
#include "sp_shade.cuh"

#define SP_BLOCK 256
#define SP_CTAS_PER_SM 2

SP_DEV void sp_stage_chunk(float4* __restrict__ dst, const DScene& sc, const GeomStream& gs, int c) {
    const int lo = __ldg(gs.chunk_off + c), hi = __ldg(gs.chunk_off + c + 1);
    const float4* __restrict__ src = gs.data + lo;
    for (int i = threadIdx.x; i < hi - lo; i += SP_BLOCK) dst[i] = __ldg(src + i);
}

__global__ void __launch_bounds__(SP_BLOCK, SP_CTAS_PER_SM)
sp_level_kernel(const __grid_constant__ DScene sc, const __grid_constant__ LevelArgs a) {
    __shared__ float4 s_geom[SP_CHUNK_VEC4];

    // ---- work items of this launch ---------------------------------------------------------
    uint32_t n_rays = 0, fan_n[SP_MAX_FAN_CLASSES];
    unsigned long long total;
#pragma unroll
    for (int c = 0; c < SP_MAX_FAN_CLASSES; ++c) fan_n[c] = 0;
    if (a.source == SP_SRC_QUEUES) {
        n_rays = min(__ldg(a.in_counts), a.in_rays.capacity);
        total = n_rays;
#pragma unroll
        for (int c = 0; c < SP_MAX_FAN_CLASSES; ++c) {
            if (c < sc.n_fan_classes) {
                fan_n[c] = min(__ldg(a.in_counts + 1 + c), a.in_fan_cap[c]);
                total += (unsigned long long)fan_n[c] * (unsigned)sc.fan_mult[c];
            }
        }
    } else {
        total = a.n_items0;
    }
    if (total == 0) return;

    const int n_chunks = sc.all.n_chunks;
    if (n_chunks == 1) {
        sp_stage_chunk(s_geom, sc, sc.all, 0);
        __syncthreads();
    }

    ShadeCtx ctx;
    ctx.sc = &sc; ctx.out = a.out; ctx.all_slot = a.all_slot; ctx.shadow_slot = a.shadow_slot;
    ctx.shadow_rays = 0;
    unsigned long long traced = 0;

    for (unsigned long long base = (unsigned long long)blockIdx.x * SP_BLOCK; base < total;
         base += (unsigned long long)gridDim.x * SP_BLOCK) {
        const unsigned long long item = base + threadIdx.x;
        bool active = item < total;
        Ray r;
        r.o = r.d = r.thr = v3(0.f); r.pix = 0; r.path = 0; r.meta = 0;

        // ---- 1. the ray of this item ---------------------------------------------------------
        if (active) {
            if (a.source == SP_SRC_CAMERA) {
                uint32_t i = (uint32_t)item;
                uint32_t sample = a.sample_begin + i / a.n_pix;
                r.pix = a.pix_begin + i % a.n_pix;
                r.path = sp_root_path(sample);
                sp_camera_ray(sc.cam, r.pix, sample, sc.seed_lo, sc.seed_hi, r.o, r.d);
                r.thr = v3(1.f);
                r.meta = sp_pack_meta(0u, 0u, 0u, SP_SRC_NONE, SP_SELF_SKIP);
            } else if (a.source == SP_SRC_USER) {
                uint32_t i = a.user_base + (uint32_t)item;
                r.pix = i;
                r.path = sp_root_path(0u);
                r.o = v3(__ldg(a.user_o + 3 * (size_t)i), __ldg(a.user_o + 3 * (size_t)i + 1), __ldg(a.user_o + 3 * (size_t)i + 2));
                r.d = v3(__ldg(a.user_d + 3 * (size_t)i), __ldg(a.user_d + 3 * (size_t)i + 1), __ldg(a.user_d + 3 * (size_t)i + 2));
                r.thr = v3(1.f);
                r.meta = sp_pack_meta(0u, 0u, 0u, SP_SRC_NONE, SP_SELF_SKIP);
            } else if (item < n_rays) {
                const uint32_t s = (uint32_t)item;
                const float4 q0 = a.in_rays.q0[s], q1 = a.in_rays.q1[s], q2 = a.in_rays.q2[s];
                r.o = xyz(q0); r.d = xyz(q1); r.thr = xyz(q2);
                r.pix = __float_as_uint(q0.w); r.path = __float_as_uint(q1.w); r.meta = __float_as_uint(q2.w);
            } else {
                unsigned long long local = item - n_rays;
                int c = 0;
#pragma unroll
                for (int k = 0; k < SP_MAX_FAN_CLASSES - 1; ++k) {
                    unsigned long long span = (unsigned long long)fan_n[k] * (unsigned)sc.fan_mult[k];
                    if (c == k && local >= span) { local -= span; c = k + 1; }
                }
                const uint32_t m = (uint32_t)sc.fan_mult[c];
                const uint32_t rec = (uint32_t)local / m, child = (uint32_t)local % m;
                const uint32_t s = a.in_fan_base[c] + rec;
                const float4 q0 = a.in_fans.q0[s], q1 = a.in_fans.q1[s], q2 = a.in_fans.q2[s];
                r.o = xyz(q0); r.thr = xyz(q2);
                r.pix = __float_as_uint(q0.w); r.meta = __float_as_uint(q2.w);
                r.path = sp_child_path(__float_as_uint(q1.w), child);
                const DCollider& sc_col = sc.colliders[meta_src(r.meta)];
                const float w_cos = sc.mats[sc.prims[sc_col.prim].material].ambient_weight;
                const float weight = sp_sample_diffuse(sc, r.o, xyz(q1), w_cos, r.pix, r.path, r.d);
                r.thr = r.thr * weight;
                active = weight > 0.f;          // zero-weight samples cannot contribute: not traced
            }
        }
        if (a.run == SP_RUN_DUMP_RAYS) {
            if (active) {
                const size_t i = (size_t)item;
                a.out_o[3 * i] = r.o.x; a.out_o[3 * i + 1] = r.o.y; a.out_o[3 * i + 2] = r.o.z;
                a.out_d[3 * i] = r.d.x; a.out_d[3 * i + 1] = r.d.y; a.out_d[3 * i + 2] = r.d.z;
            }
            continue;
        }

        // ---- 2. nearest hit over all colliders ------------------------------------------------------
        HitRec hit; hit.t = SP_INF; hit.id = -1; hit.orient = 0;
        const uint32_t src = meta_src(r.meta), mode = meta_mode(r.meta);
        int2 where = make_int2(-1, -1);
        bool need_test = active;
        if (active && src != SP_SRC_NONE) {
            if (mode == SP_SELF_ZERO) {
                // the ray dives back into the surface it starts on: the reference re-hits it after
                // ~1e-6 (its nudge); here that is an immediate hit at t = 0
                const DCollider& c0 = sc.colliders[src];
                float3 Nc = to_f3(sp_collider_normal<float>(c0.type, c0.p, from_f3<float>(r.o)));
                hit.t = 0.f; hit.id = (int)src; hit.orient = dot(r.d, Nc) < 0.f ? 1 : -1;
                need_test = false;
            } else {
                where = __ldg(a.all_slot + src);
            }
        }
        for (int c = 0; c < n_chunks; ++c) {
            if (n_chunks > 1) {
                __syncthreads();
                sp_stage_chunk(s_geom, sc, sc.all, c);
                __syncthreads();
            }
            if (need_test) {
                SelfSlot self; self.sphere = self.plane = self.cuboid = self.tri = -1; self.mode = mode;
                if (where.x == c) {
                    const int ty = where.y >> 28, li = where.y & 0x0FFFFFFF;
                    if (ty == 0) self.sphere = li; else if (ty == 1) self.plane = li;
                    else if (ty == 2) self.cuboid = li; else self.tri = li;
                }
                ChunkBest best; best.t = hit.t; best.idx = -1; best.orient = 0;
                sp_intersect_chunk(s_geom, r.o, r.d, self, best);
                if (best.idx >= 0) { hit.t = best.t; hit.orient = best.orient; hit.id = sp_chunk_id(s_geom, best.idx); }
            }
        }
        if (!active) continue;
        traced += 1;

        if (a.level == 0) {
            const size_t oi = (a.source == SP_SRC_USER) ? (size_t)a.user_base + (size_t)item : (size_t)item;
            if (a.out_hit) a.out_hit[oi] = hit.id;
            if (a.out_t) a.out_t[oi] = hit.t;
        }
        if (a.run == SP_RUN_DISTANCES || hit.id < 0) continue;

        // ---- 3. shade, accumulate, emit children ------------------------------------------------------
        const float3 add = sp_shade(ctx, r, hit);
        float* px = reinterpret_cast<float*>(a.accum + r.pix);
        if (add.x != 0.f) atomicAdd(px, add.x);
        if (add.y != 0.f) atomicAdd(px + 1, add.y);
        if (add.z != 0.f) atomicAdd(px + 2, add.z);
    }

    // ---- counters: one atomic per warp -------------------------------------------------------------
    unsigned long long sh = ctx.shadow_rays;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        traced += __shfl_down_sync(0xffffffffu, traced, o);
        sh += __shfl_down_sync(0xffffffffu, sh, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (traced) atomicAdd(&a.out.stats->rays[a.level], traced);
        if (sh) atomicAdd(&a.out.stats->shadow_rays, sh);
    }
}

// ---- frame resolve: average, sRGB OETF, per-pixel max normalisation, truncation to uint8 ---------
// scene.py:118-140 + colour_functions.py:4-18.  Double precision: the output is quantised by
// truncation, so the transfer curve must not wobble around integer boundaries.
__global__ void __launch_bounds__(256) sp_resolve_kernel(const ResolveArgs a) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= a.n_pix) return;
    const float4 acc = a.accum[i];
    const float lin[3] = {(float)((double)acc.x / a.spp), (float)((double)acc.y / a.spp),
                          (float)((double)acc.z / a.spp)};
    if (a.out_linear) {
        a.out_linear[i] = lin[0];
        a.out_linear[(size_t)a.n_pix + i] = lin[1];
        a.out_linear[2 * (size_t)a.n_pix + i] = lin[2];
    }
    if (!a.out_srgb8) return;
    double enc[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double x = (double)lin[c];
        enc[c] = (x <= 0.00304) ? 12.92 * x : 1.055 * pow(x, 1.0 / 2.4) - 0.055;
    }
    // np.amax propagates NaN; comparisons with NaN are false (no rescale)
    double peak = fmax(fmax(enc[0], enc[1]), enc[2]);
    if (enc[0] != enc[0] || enc[1] != enc[1] || enc[2] != enc[2]) peak = enc[0] + enc[1] + enc[2];
    peak += 0.00001;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double e = enc[c];
        if (peak > 1.0) e = e * 1.0 / peak;
        e = fmin(fmax(e, 0.0), 1.0);                       // np.clip
        const double s = 255.0 * e;
        a.out_srgb8[3 * (size_t)i + c] = (s == s) ? (uint8_t)s : (uint8_t)0;
    }
}

// ---- roofline denominators ----------------------------------------------------------------------
__global__ void __launch_bounds__(256) sp_ffma_kernel(float* out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = (float)(threadIdx.x + k);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = fmaf(x[k], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += x[k];
    if (s == 12345.678f) out[0] = s;                      // keep the chain alive
}

__global__ void __launch_bounds__(256) sp_copy_kernel(const float4* __restrict__ src, float4* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) dst[i] = src[i];
}

// =================================================================================================
// launchers
// =================================================================================================
int sp_level_grid(int device) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    int per_sm = SP_CTAS_PER_SM;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sp_level_kernel, SP_BLOCK, 0) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    return sms * per_sm;
}

cudaError_t sp_launch_level(const DScene& sc, const LevelArgs& a, int grid, cudaStream_t st) {
    sp_level_kernel<<<grid, SP_BLOCK, 0, st>>>(sc, a);
    return cudaGetLastError();
}

cudaError_t sp_launch_resolve(const ResolveArgs& a, cudaStream_t st) {
    if (a.n_pix == 0) return cudaSuccess;
    sp_resolve_kernel<<<(a.n_pix + 255u) / 256u, 256, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t sp_upload_decode_tables(const float* plain256, const float* linear256) {
    cudaError_t e = cudaMemcpyToSymbol(c_decode, plain256, 256 * sizeof(float), 0);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_decode, linear256, 256 * sizeof(float), 256 * sizeof(float));
}

cudaError_t sp_bench_ffma(double* tflops, cudaStream_t st) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    float* out = nullptr;
    cudaError_t e = cudaMalloc(&out, 4);
    if (e != cudaSuccess) return e;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    const int grid = sms * 8, iters = 1 << 14;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(t0, st);
        sp_ffma_kernel<<<grid, 256, 0, st>>>(out, iters, 1.000001f, 1e-7f);
        cudaEventRecord(t1, st);
        e = cudaEventSynchronize(t1);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        const double flop = 2.0 * 16.0 * iters * 256.0 * grid;
        if (rep > 0 && ms > 0.f) best = fmax(best, flop / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    cudaFree(out);
    *tflops = best;
    return e;
}

cudaError_t sp_bench_copy(double* gbs, cudaStream_t st) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t n = (size_t)1 << 26;                         // 64 Mi float4 = 1 GiB per buffer
    float4 *a = nullptr, *b = nullptr;
    cudaError_t e = cudaMalloc(&a, n * sizeof(float4));
    if (e != cudaSuccess) return e;
    e = cudaMalloc(&b, n * sizeof(float4));
    if (e != cudaSuccess) { cudaFree(a); return e; }
    cudaMemsetAsync(a, 0, n * sizeof(float4), st);
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(t0, st);
        sp_copy_kernel<<<sms * 16, 256, 0, st>>>(a, b, n);
        cudaEventRecord(t1, st);
        e = cudaEventSynchronize(t1);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        if (rep > 0 && ms > 0.f) best = fmax(best, 2.0 * n * sizeof(float4) / (ms * 1e-3) / 1e9);
    }
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    cudaFree(a); cudaFree(b);
    *gbs = best;
    return e;
}
