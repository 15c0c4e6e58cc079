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
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "sp_launch.h"
#include "sp_sampling.cuh"
#include "sp_shade.cuh"

#ifndef SP_BLOCK
#define SP_BLOCK 256
#endif
// resident CTAs per SM the register allocation aims for: the lean Monte-Carlo variant fits 64
// registers (4 CTAs) with a handful of spills; the textured / glossy / BVH variants spill more at 64 registers
// but still gain 4-12 % from the fourth CTA (measured: example2 +4 %, example4 +5 %, stress scene +12 %)
#ifndef SP_CTAS_MC
#define SP_CTAS_MC 4
#endif
#ifndef SP_CTAS_FULL
#define SP_CTAS_FULL 4
#endif
#ifndef SP_LEVEL_DYNAMIC
#define SP_LEVEL_DYNAMIC 1
#endif
#define SP_CTAS_PER_SM(FEAT) ((((FEAT) & (SP_F_TEX | SP_F_GLOSSY | SP_F_THIN | SP_F_SKY | SP_F_BVH)) == 0u) ? SP_CTAS_MC : SP_CTAS_FULL)

SP_DEV void sp_stage_chunk(float4* __restrict__ dst, const DScene& sc, const GeomStream& gs, int c) {
    const int lo = __ldg(gs.chunk_off + c), hi = __ldg(gs.chunk_off + c + 1);
    const float4* __restrict__ src = gs.data + lo;
    for (int i = threadIdx.x; i < hi - lo; i += SP_BLOCK) dst[i] = __ldg(src + i);
}

// Per-CTA exchange area of one iteration (SP_BATCH rays, SP_RPT per thread): the rays and their
// hits are parked here after the intersection phase, regrouped by the material kind they hit
// (most expensive kind first), and picked up again in 32-ray chunks by whichever warp is free, so
// that the lanes of a warp shade the same material and the warps of a CTA finish together (the
// fused kernel would otherwise run Diffuse / Refractive / Emissive code with a third of its lanes
// each, and the warp that drew the glass hits would keep the other seven waiting).
#ifndef SP_RPT
#define SP_RPT 2
#endif
#define SP_BATCH (SP_BLOCK * SP_RPT)
#define SP_STATE_WORDS 16
#define SP_N_BINS 7              // the six material kinds + "nothing to shade"
#define SP_N_QUEUES (1 + SP_MAX_FAN_CLASSES)
#define SP_N_WARPS (SP_BLOCK / 32)
#define SP_N_VWARPS (SP_N_WARPS * SP_RPT)
// material kind -> shading bin, dearest first: Refractive, Glossy, ThinFilm, Diffuse, SkyBox, Emissive
#define SP_BIN_OF_KIND(kind) ((0x453201u >> (4u * (kind))) & 15u)
struct IterCounters {
    uint32_t bin_cnt[8];                           // rays per bin
    uint32_t queue_cnt[8];                         // records this iteration appends to each output queue
    uint32_t queue_base[8];                        // the CTA's reservation in each output queue
    uint32_t arrived;                              // (pass, warp)s that have counted; the last one reserves
    uint32_t next_chunk;                           // next 32-ray shading chunk up for grabs
    uint32_t pad[6];
};
struct IterShared {
    uint32_t state[SP_STATE_WORDS][SP_BATCH];      // SoA: o d thr pix path meta t (id|orient) ray_slot fan_slot
    uint16_t list[SP_N_BINS - 1][SP_BATCH];        // per shading bin: batch slots of the rays that hit that kind
    IterCounters cnt[2];                           // alternate between iterations: the idle set is cleared
                                                   // while the other is in use, which saves a barrier
};

// Development aid: build with -DSP_PHASE_TIMING and run with SIGHTPY_PHASE_TIMING=1 to get the share of
// warp-cycles each phase of the kernel takes (clock64() at the phase boundaries, printed by sp_api.cu).
#ifdef SP_PHASE_TIMING
#define SP_TICK(k) do { const long long now__ = clock64(); phase_acc[k] += (unsigned long long)(now__ - tick__); tick__ = now__; } while (0)
#else
#define SP_TICK(k) do { } while (0)
#endif

// The ray of work item `item` of a level launch (camera / caller ray at level 0; a queued ray record, or child
// `item % mult` of a queued fan record, afterwards).  Returns false for items that carry nothing (dead records,
// zero-weight samples, texels of an edge tile outside the frame).  Deterministic in (scene, args, item): the trace
// kernel of BVH scenes and the level kernel both call it and see the same ray.
template <uint32_t FEAT>
SP_DEV bool sp_item_ray(const DScene& sc, const LevelArgs& a, uint32_t item, uint32_t n_rays, const uint32_t* fan_n, Ray& r) {
    bool active = true;
    if ((FEAT & SP_F_LEVEL0) && a.source == SP_SRC_CAMERA) {
        const uint32_t i = (uint32_t)item;
        const uint32_t si = a.n_pix > 1u ? (uint32_t)__umul64hi((unsigned long long)i, a.n_pix_magic) : i;   // i / n_pix
        const uint32_t sample = a.sample_begin + si;
        r.pix = a.pix_begin + (i - si * a.n_pix);
        if (a.tiles) {                                // texel of a tile list -> pixel of the frame
            const uint32_t ts = a.tile_shift, local = r.pix & ((1u << (2u * ts)) - 1u);
            const uint32_t tile = __ldg(a.tiles + (r.pix >> (2u * ts)));
            const uint32_t px = ((tile % a.tiles_x) << ts) + (local & ((1u << ts) - 1u));
            const uint32_t py = ((tile / a.tiles_x) << ts) + (local >> ts);
            active = px < (uint32_t)sc.cam.W && py < (uint32_t)sc.cam.H;
            r.pix = active ? py * (uint32_t)sc.cam.W + px : 0u;
        }
        if (active) {
            r.path = sp_root_path(sample);
            sp_camera_ray(sc.cam, r.pix, sample, sc.philox_keys, r.o, r.d);
            r.thr = v3(1.f);
            r.meta = sp_pack_meta(0u, 0u, 0u, SP_SRC_NONE, SP_SELF_SKIP);
        }
    } else if ((FEAT & SP_F_LEVEL0) && a.source == SP_SRC_USER) {
        uint32_t i = a.user_base + (uint32_t)item;
        r.pix = i;
        r.path = sp_root_path(0u);
        r.o = v3(__ldg(a.user_o + 3 * (size_t)i), __ldg(a.user_o + 3 * (size_t)i + 1), __ldg(a.user_o + 3 * (size_t)i + 2));
        r.d = v3(__ldg(a.user_d + 3 * (size_t)i), __ldg(a.user_d + 3 * (size_t)i + 1), __ldg(a.user_d + 3 * (size_t)i + 2));
        r.thr = v3(1.f);
        r.meta = sp_pack_meta(0u, 0u, 0u, SP_SRC_NONE, SP_SELF_SKIP);
    } else if (!(FEAT & SP_F_QUEUES)) {
        active = false;
    } else if (item < n_rays) {
        const uint32_t s = (uint32_t)item;
        // all three vectors at once: one memory round trip (a dead record's q0 / q1 are simply ignored)
        const float4 q2 = a.in_rays.q2[s], q0 = a.in_rays.q0[s], q1 = a.in_rays.q1[s];
        r.meta = __float_as_uint(q2.w);
        if (r.meta == SP_META_DEAD) {
            active = false;
        } else {
            r.o = xyz(q0); r.d = xyz(q1); r.thr = xyz(q2);
            r.pix = __float_as_uint(q0.w); r.path = __float_as_uint(q1.w);
        }
    } else if (!(FEAT & SP_F_DIFFUSE)) {
        active = false;
    } else {
        uint32_t local = item - n_rays;
        int c = 0;
#pragma unroll
        for (int k = 0; k < SP_MAX_FAN_CLASSES - 1; ++k) {
            const uint32_t span = fan_n[k] * (uint32_t)sc.fan_mult[k];
            if (c == k && local >= span) { local -= span; c = k + 1; }
        }
        const uint32_t m = (uint32_t)sc.fan_mult[c];
        const uint32_t rec = (m == 1u) ? local : (uint32_t)__umul64hi((unsigned long long)local, sc.fan_magic[c]);
        const uint32_t child = local - rec * m;
        const uint32_t s = a.in_fan_base[c] + rec;
        const float4 q2 = a.in_fans.q2[s], q0 = a.in_fans.q0[s], q1 = a.in_fans.q1[s];
        r.meta = __float_as_uint(q2.w);
        if (r.meta == SP_META_DEAD) {
            active = false;
        } else {
            r.o = xyz(q0); r.thr = xyz(q2);
            r.pix = __float_as_uint(q0.w);
            r.path = sp_child_path(__float_as_uint(q1.w), child);
            const float w_cos = __ldg(&sc.col_info[meta_src(r.meta)].w_cos);
            const float weight = sp_sample_diffuse(sc, r.o, xyz(q1), w_cos, r.pix, r.path, r.d);
            r.thr = r.thr * weight;
            active = weight > 0.f;          // zero-weight samples cannot contribute: not traced
        }
    }
    return active;
}

template <uint32_t FEAT>
__global__ void __launch_bounds__(SP_BLOCK, SP_CTAS_PER_SM(FEAT))
sp_level_kernel(const __grid_constant__ DScene sc, const __grid_constant__ LevelArgs a) {
    extern __shared__ float4 s_geom[];                 // sized by the host to the scene's largest chunk
    __shared__ IterShared sh;

    // a queue of an earlier level overflowed: its records are incomplete, the host discards the chunk
    if (*reinterpret_cast<volatile const unsigned int*>(&a.out.stats->overflow) & 0xFFFFu) return;

    // ---- work items of this launch ---------------------------------------------------------
    uint32_t n_rays = 0, fan_n[SP_MAX_FAN_CLASSES];
    uint32_t total;                                    // < 2^32 work items per launch (checked by the host)
#pragma unroll
    for (int c = 0; c < SP_MAX_FAN_CLASSES; ++c) fan_n[c] = 0;
    if ((FEAT & SP_F_QUEUES) && a.source == SP_SRC_QUEUES) {
        n_rays = min(__ldg(a.in_counts), a.in_rays.capacity);
        total = n_rays;
#pragma unroll
        for (int c = 0; c < SP_MAX_FAN_CLASSES; ++c) {
            if (c < sc.n_fan_classes) {
                fan_n[c] = min(__ldg(a.in_counts + 1 + c), a.in_fan_cap[c]);
                total += fan_n[c] * (uint32_t)sc.fan_mult[c];
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

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    ShadeCtx ctx;
    ctx.sc = &sc; ctx.out = &a.out; ctx.shadow_slot = a.shadow_slot;
    ctx.shq = a.shq; ctx.shq_cap = a.shq_cap; ctx.shq_count = a.shq_count;
    ctx.shadow_rays = 0;
    unsigned long long traced = 0;

    __shared__ float s_lin_lut[(FEAT & SP_F_TEX) ? 256 : 1];
    if (FEAT & SP_F_TEX)
        for (uint32_t i = tid; i < 256u; i += SP_BLOCK) s_lin_lut[i] = c_decode[SP_DECODE_LINEAR][i];
    ctx.lin_lut = s_lin_lut;
    if (tid < sizeof(sh.cnt) / sizeof(uint32_t)) reinterpret_cast<uint32_t*>(sh.cnt)[tid] = 0u;
    // Work distribution of full runs: CTAs draw their 512-item iterations from a counter (first word of the work
    // block behind the level's queue counts, zeroed by the host) instead of striding over them; rays differ in
    // cost, and a static split leaves SMs idle at the end of every launch.  The draw for iteration i + 1 is issued
    // at the top of iteration i and handed to the CTA through shared memory behind barrier (C).
    __shared__ uint32_t s_work[2];
    const bool dynamic = SP_LEVEL_DYNAMIC && a.run == SP_RUN_FULL;
    uint32_t* const work = const_cast<uint32_t*>(a.in_counts) + SP_COUNTS_PER_LEVEL / 2;
    if (dynamic && tid == 0) s_work[0] = atomicAdd(work, (uint32_t)SP_BATCH);
    __syncthreads();
#ifdef SP_PHASE_TIMING
    unsigned long long phase_acc[6] = {0, 0, 0, 0, 0, 0};
    long long tick__ = clock64();
    unsigned long long gt0__; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt0__));
    const long long ck0__ = tick__;
#endif
    uint32_t parity = 0;
    for (unsigned long long base64 = dynamic ? (unsigned long long)s_work[0] : (unsigned long long)blockIdx.x * SP_BATCH;
         base64 < total;
         base64 = dynamic ? (unsigned long long)s_work[parity ^ 1u] : base64 + (unsigned long long)gridDim.x * SP_BATCH, parity ^= 1u) {
        IterCounters& cn = sh.cnt[parity];
        const uint32_t base = (uint32_t)base64;
        uint32_t drawn = 0;
        if (dynamic && tid == 0) drawn = atomicAdd(work, (uint32_t)SP_BATCH);
        // the pass loop is deliberately not unrolled: one copy of the generate/intersect code in the
        // instruction cache
#pragma unroll 1
        for (int pass = 0; pass < SP_RPT; ++pass) {
        const uint32_t slot = (uint32_t)pass * SP_BLOCK + tid;
        const uint32_t item = base + slot;
        bool active = (unsigned long long)base + slot < total;
        Ray r;
        r.o = r.d = r.thr = v3(0.f); r.pix = 0; r.path = 0; r.meta = 0;

        // ---- 1. the ray of this item ---------------------------------------------------------
        if (active) active = sp_item_ray<FEAT>(sc, a, item, n_rays, fan_n, r);
        if ((FEAT & SP_F_LEVEL0) && a.run == SP_RUN_DUMP_RAYS) {
            if (active) {
                const size_t i = (size_t)item;
                a.out_o[3 * i] = r.o.x; a.out_o[3 * i + 1] = r.o.y; a.out_o[3 * i + 2] = r.o.z;
                a.out_d[3 * i] = r.d.x; a.out_d[3 * i + 1] = r.d.y; a.out_d[3 * i + 2] = r.d.z;
            }
            continue;
        }

        SP_TICK(0);
        // ---- 2. nearest hit over all colliders ------------------------------------------------------
        HitRec hit; hit.t = SP_INF; hit.id = -1; hit.orient = 0;
        {
            const uint32_t src = meta_src(r.meta), mode = meta_mode(r.meta);
            uint32_t where = 0xFFFFFFFFu;                  // chunk | stream type | local index of the source collider
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
                    where = __ldg(&sc.col_info[src].slot);
                }
            }
            if ((FEAT & SP_F_BVH) && need_test && a.hits && total <= a.hits_cap) {
                // found ahead of this launch by sp_trace_kernel (full occupancy, no barrier behind the slowest traversal)
                const float2 h = a.hits[item];
                const uint32_t code = __float_as_uint(h.y);
                hit.t = h.x;
                hit.id = (code & 0x7FFFFFFFu) == 0x7FFFFFFFu ? -1 : (int)(code & 0x7FFFFFFFu);
                hit.orient = (code & 0x80000000u) ? 1 : -1;
                need_test = false;
            }
            for (int c = 0; c < n_chunks; ++c) {
                if (n_chunks > 1) {
                    __syncthreads();
                    sp_stage_chunk(s_geom, sc, sc.all, c);
                    __syncthreads();
                }
                if (need_test) {
                    SelfSlot self; self.sphere = self.plane = self.cuboid = self.tri = self.aa = -1; self.mode = mode;
                    if ((where >> 24) == (uint32_t)c) {
                        const int ty = (int)((where >> 20) & 15u), li = (int)(where & 0xFFFFFu);
                        if (ty == 0) self.sphere = li; else if (ty == 1) self.plane = li;
                        else if (ty == 2) self.cuboid = li; else if (ty == 3) self.tri = li; else self.aa = li;
                    }
                    ChunkBest best; best.t = hit.t; best.idx = -1; best.orient = 0;
                    sp_intersect_chunk(s_geom, r.o, r.d, self, best);
                    if (best.idx >= 0) { hit.t = best.t; hit.orient = best.orient; hit.id = sp_chunk_id(s_geom, best.idx); }
                }
            }
            if ((FEAT & SP_F_BVH) && need_test) {
                ChunkBest best; best.t = hit.t; best.idx = -1; best.orient = 0;
                sp_bvh_nearest(sc.bvh, r.o, r.d, src == SP_SRC_NONE ? -1 : (int)src, mode, false, -SP_INF, best);
                if (best.idx >= 0) { hit.t = best.t; hit.orient = best.orient; hit.id = best.idx; }
            }
        }
        if (active) {
            traced += 1;
            if ((FEAT & SP_F_LEVEL0) && a.level == 0 && (a.out_hit || a.out_t || a.out_n)) {       // per-ray outputs (sp_trace, sp_aovs)
                const size_t oi = (a.source == SP_SRC_USER) ? (size_t)a.user_base + (size_t)item : (size_t)item;
                if (a.out_hit) a.out_hit[oi] = hit.id;
                if (a.out_t) a.out_t[oi] = hit.t;
                if (a.out_n) {                           // collider normal facing the ray (collider.get_Normal x orientation)
                    float3 n = v3(0.f);
                    if (hit.id >= 0) {
                        const DCollider& c0 = sc.colliders[hit.id];
                        n = to_f3(sp_collider_normal<float>(c0.type, c0.p, from_f3<float>(fma3(r.d, hit.t, r.o)))) * (float)hit.orient;
                    }
                    a.out_n[3 * oi] = n.x; a.out_n[3 * oi + 1] = n.y; a.out_n[3 * oi + 2] = n.z;
                }
            }
        }
        if ((FEAT & SP_F_LEVEL0) && a.run == SP_RUN_DISTANCES) continue;

        SP_TICK(1);
        // ---- 3. park the ray; what the hit will emit; per-warp counts of bins and queue records ---------
        sh.state[0][slot] = __float_as_uint(r.o.x); sh.state[1][slot] = __float_as_uint(r.o.y); sh.state[2][slot] = __float_as_uint(r.o.z);
        sh.state[3][slot] = __float_as_uint(r.d.x); sh.state[4][slot] = __float_as_uint(r.d.y); sh.state[5][slot] = __float_as_uint(r.d.z);
        sh.state[6][slot] = __float_as_uint(r.thr.x); sh.state[7][slot] = __float_as_uint(r.thr.y); sh.state[8][slot] = __float_as_uint(r.thr.z);
        sh.state[9][slot] = r.pix; sh.state[10][slot] = r.path; sh.state[11][slot] = r.meta;
        sh.state[12][slot] = __float_as_uint(hit.t);
        sh.state[13][slot] = (uint32_t)hit.id | (hit.orient > 0 ? 0x80000000u : 0u);
        int bin = SP_N_BINS - 1, n_ray = 0, fan_class = -1;
        if (active && hit.id >= 0) {
            const float4 raw = __ldg(reinterpret_cast<const float4*>(sc.col_info + hit.id));
            const DColInfo ci = *reinterpret_cast<const DColInfo*>(&raw);
            bin = (int)SP_BIN_OF_KIND(ci.kind);
            sp_child_needs(ci, meta_depth(r.meta), meta_dr(r.meta), n_ray, fan_class);
            // a Diffuse hit past its bounce budget and a Refractive / ThinFilm hit past max_ray_depth are
            // black (diffuse.py:123-124, refractive.py:38): nothing to shade, nothing to emit
            if (n_ray == 0 && fan_class < 0 &&
                (ci.kind == SP_MAT_DIFFUSE || ci.kind == SP_MAT_REFRACTIVE || ci.kind == SP_MAT_THINFILM))
                bin = SP_N_BINS - 1;
        }
        // position inside the bin, and inside this iteration's appends to the output queues: one
        // shared-memory atomic per warp and distinct bin / queue
        if (bin < SP_N_BINS - 1) {
            const uint32_t peers = __match_any_sync(__activemask(), bin);
            uint32_t pos = 0;
            const int leader = __ffs(peers) - 1;
            if ((int)lane == leader) pos = atomicAdd(&cn.bin_cnt[bin], (uint32_t)__popc(peers));
            pos = __shfl_sync(peers, pos, leader) + __popc(peers & lt_mask);
            SP_ASSERT(a.out.stats, pos < (uint32_t)SP_BATCH, SP_CHK_BIN);
            sh.list[bin][pos] = (uint16_t)slot;
        }
        {
            const uint32_t b0 = __ballot_sync(0xffffffffu, n_ray & 1), b1 = __ballot_sync(0xffffffffu, n_ray & 2);
            const uint32_t warp_rays = __popc(b0) + 2u * __popc(b1);
            uint32_t ray_off = 0;
            if (warp_rays) {
                if (lane == 0) ray_off = atomicAdd(&cn.queue_cnt[0], warp_rays);
                ray_off = __shfl_sync(0xffffffffu, ray_off, 0) + __popc(b0 & lt_mask) + 2u * __popc(b1 & lt_mask);
            }
            sh.state[14][slot] = (uint32_t)n_ray | (ray_off << 2);
            uint32_t fan_word = SP_SLOT_NONE;
            if (FEAT & SP_F_DIFFUSE) {
                const uint32_t fan_any = __ballot_sync(0xffffffffu, fan_class >= 0);
                if (fan_any) {
#pragma unroll
                    for (int c = 0; c < SP_MAX_FAN_CLASSES; ++c) {
                        const uint32_t bc = __ballot_sync(0xffffffffu, fan_class == c);
                        if (bc) {
                            uint32_t off = 0;
                            if (lane == 0) off = atomicAdd(&cn.queue_cnt[1 + c], (uint32_t)__popc(bc));
                            off = __shfl_sync(0xffffffffu, off, 0) + __popc(bc & lt_mask);
                            if (fan_class == c) fan_word = ((uint32_t)c << 28) | off;
                        }
                    }
                }
            }
            sh.state[15][slot] = fan_word;
        }
        // the last (pass, warp) to get here knows the totals: it reserves the CTA's slots in the
        // output queues (one global atomic per queue that receives something)
        __syncwarp();
        uint32_t ticket = 0;
        if (lane == 0) { __threadfence_block(); ticket = atomicAdd(&cn.arrived, 1u); }
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket == SP_N_VWARPS - 1u && lane < SP_N_QUEUES) {
            __threadfence_block();
            const uint32_t want = cn.queue_cnt[lane];
            const uint32_t cap = (lane == 0) ? a.out.rays.capacity : a.out.fan_cap[lane - 1];
            uint32_t first = SP_SLOT_NONE;
            if (want > 0) {
                first = atomicAdd(a.out.counts + lane, want);
                if (first + want > cap || first + want < first) { first = SP_SLOT_NONE; atomicOr(&a.out.stats->overflow, 1u); }
                else if (lane > 0) first += a.out.fan_base[lane - 1];
            }
            cn.queue_base[lane] = first;
        }
        }   // pass
        if ((FEAT & SP_F_LEVEL0) && a.run != SP_RUN_FULL) continue;
        SP_TICK(2);
        __syncthreads();                                                            // (A) everything parked
        SP_TICK(3);
        // clear the other counter set: last touched before barrier (C) of the previous iteration,
        // next touched after barrier (C) of this one
        if (tid < sizeof(IterCounters) / sizeof(uint32_t)) reinterpret_cast<uint32_t*>(&sh.cnt[parity ^ 1u])[tid] = 0u;

        // ---- 4. shade bin by bin (32-ray chunks handed out on demand), accumulate, write children -----------
        uint32_t bin_chunks_end[SP_N_BINS - 1];         // running number of chunks up to and including each bin
        {
            uint32_t run = 0;
#pragma unroll
            for (int b = 0; b < SP_N_BINS - 1; ++b) { run += (cn.bin_cnt[b] + 31u) >> 5; bin_chunks_end[b] = run; }
        }
        while (true) {
            uint32_t chunk = 0;
            if (lane == 0) chunk = atomicAdd(&cn.next_chunk, 1u);
            chunk = __shfl_sync(0xffffffffu, chunk, 0);
            if (chunk >= bin_chunks_end[SP_N_BINS - 2]) break;
            int b = 0;
#pragma unroll
            for (int q = 0; q < SP_N_BINS - 2; ++q) b += (chunk >= bin_chunks_end[q]) ? 1 : 0;
            uint32_t chunk_in_bin = chunk;
#pragma unroll
            for (int q = 0; q < SP_N_BINS - 2; ++q) if (b == q + 1) chunk_in_bin = chunk - bin_chunks_end[q];
            const uint32_t k = chunk_in_bin * 32u + lane;
            if (k < cn.bin_cnt[b]) {
                const uint32_t j = sh.list[b][k];

                Ray s;
                s.o = v3(__uint_as_float(sh.state[0][j]), __uint_as_float(sh.state[1][j]), __uint_as_float(sh.state[2][j]));
                s.d = v3(__uint_as_float(sh.state[3][j]), __uint_as_float(sh.state[4][j]), __uint_as_float(sh.state[5][j]));
                s.thr = v3(__uint_as_float(sh.state[6][j]), __uint_as_float(sh.state[7][j]), __uint_as_float(sh.state[8][j]));
                s.pix = sh.state[9][j]; s.path = sh.state[10][j]; s.meta = sh.state[11][j];
                HitRec h;
                h.t = __uint_as_float(sh.state[12][j]);
                const uint32_t packed = sh.state[13][j];
                h.id = (int)(packed & 0x7FFFFFFFu); h.orient = (packed & 0x80000000u) ? 1 : -1;
                const uint32_t rs = sh.state[14][j], fs = sh.state[15][j];
                const uint32_t need_ray = rs & 3u;
                const uint32_t rbase = cn.queue_base[0];
                ctx.ray_slot = (need_ray && rbase != SP_SLOT_NONE) ? rbase + (rs >> 2) : SP_SLOT_NONE;
                ctx.ray_slot1 = ctx.ray_slot == SP_SLOT_NONE ? SP_SLOT_NONE : ctx.ray_slot + 1u;
                ctx.ray_used = 0u;
                ctx.fan_slot = SP_SLOT_NONE;
                if (fs != SP_SLOT_NONE) {
                    const uint32_t fbase = cn.queue_base[1 + (fs >> 28)];
                    if (fbase != SP_SLOT_NONE) ctx.fan_slot = fbase + (fs & 0x0FFFFFFFu);
                }
                const float3 add = sp_shade<FEAT>(ctx, s, h);
                sp_accum_add(a.accum + s.pix, add);
                // reserved but unused slots become dead records
                if (ctx.ray_slot != SP_SLOT_NONE)
                    for (uint32_t q = ctx.ray_used; q < need_ray; ++q) sp_write_dead(a.out.rays, ctx.ray_slot + q);
                if (ctx.fan_slot != SP_SLOT_NONE) sp_write_dead(a.out.fans, ctx.fan_slot);
            }
        }
        SP_TICK(4);
        if (dynamic && tid == 0) s_work[parity ^ 1u] = drawn;
        __syncthreads();                                                            // (C) the exchange area is free again
        SP_TICK(5);
    }

    // ---- counters: one atomic per warp -------------------------------------------------------------
    unsigned long long shr = ctx.shadow_rays;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        traced += __shfl_down_sync(0xffffffffu, traced, o);
        shr += __shfl_down_sync(0xffffffffu, shr, o);
    }
#ifdef SP_PHASE_TIMING
    if (lane == 0)
        for (int k = 0; k < 6; ++k) atomicAdd(&a.out.stats->phase_cycles[k], phase_acc[k]);
    if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) {
        unsigned long long gt1__; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt1__));
        printf("[pt] level %d cta %d start %llu dur_ns %llu cycles %lld\n", a.level, (int)blockIdx.x, gt0__, gt1__ - gt0__, clock64() - ck0__);
    }
#endif
    if (lane == 0) {
        if (traced) atomicAdd(&a.out.stats->rays[a.level], traced);
        if (shr) atomicAdd(&a.out.stats->shadow_rays, shr);
    }
}

// ---- nearest hits ahead of the level launch (scenes behind a BVH) -----------------------------------------------------
// BVH traversal lengths differ wildly from ray to ray.  Inside sp_level_kernel every 512-ray batch waits at a barrier
// for its slowest ray, and a warp that walks 32 rays in lock-step runs with the lanes of its longest ray only (ncu,
// stress scene: 8 of 32 lanes per instruction).  This kernel does the level's "generate + intersect" half on its
// own, with *persistent lanes* (Aila & Laine, Understanding the efficiency of ray traversal on GPUs):
//   * a warp prepares 32 rays at a time with full lanes — sp_item_ray (camera ray / queue record / sampled fan child)
//     and the walk over the staged chunk of scene-sized colliders, the same calls as the level kernel — and parks
//     them in a warp-private shared-memory queue;
//   * every lane traverses the BVH for one ray of its own; a lane whose ray is finished writes (t, id | orientation)
//     for the work item and pops the next prepared ray, so the lanes stay busy however long their neighbours' rays take;
//   * one round of the loop = box nodes until the lane holds a leaf, then that leaf's colliders ("while-while").
// The level launch that follows reads the hits instead of intersecting, and only parks and shades.
#define SPT_BLOCK 256
#define SPT_WARPS (SPT_BLOCK / 32)
#define SPT_BATCH 8                  // groups of 32 items a warp draws per atomic
#ifndef SPT_CTAS
#define SPT_CTAS 4                   // resident CTAs per SM the register allocation aims for
#endif
#define SPT_REC_WORDS 12             // o d t code src mode item -
struct TraceShared { uint32_t q[SPT_WARPS][SPT_REC_WORDS][32]; };

// Prepare the rays of items [first, first + 32) (those below `total`) and park the ones that need a BVH traversal in the
// warp's queue; returns how many were parked.  Rays that carry nothing, and rays that re-hit their own surface at
// t = 0, need no hit record: the level kernel answers them itself.
template <uint32_t FEAT>
__device__ __noinline__ uint32_t sp_trace_prepare(const DScene* scp, const LevelArgs* ap, const float4* s_geom, uint32_t* q,
                                                  uint32_t first, uint32_t total, uint32_t n_rays, const uint32_t* fan_n_in) {
    const DScene& sc = *scp;
    const LevelArgs& a = *ap;
    const uint32_t lane = threadIdx.x & 31u, item = first + lane;
    uint32_t fan_n[SP_MAX_FAN_CLASSES];
#pragma unroll
    for (int c = 0; c < SP_MAX_FAN_CLASSES; ++c) fan_n[c] = fan_n_in[c];
    Ray r;
    r.o = r.d = r.thr = v3(0.f); r.pix = 0; r.path = 0; r.meta = 0;
    bool active = item < total;
    if (active) active = sp_item_ray<FEAT>(sc, a, item, n_rays, fan_n, r);
    const uint32_t src = meta_src(r.meta), mode = meta_mode(r.meta);
    if (src != SP_SRC_NONE && mode == SP_SELF_ZERO) active = false;
    ChunkBest best; best.t = SP_INF; best.idx = -1; best.orient = 0;
    int hit_id = -1;
    if (active) {
        SelfSlot self; self.sphere = self.plane = self.cuboid = self.tri = self.aa = -1; self.mode = mode;
        if (src != SP_SRC_NONE) {
            const uint32_t where = __ldg(&sc.col_info[src].slot);
            if ((where >> 24) == 0u) {
                const int ty = (int)((where >> 20) & 15u), li = (int)(where & 0xFFFFFu);
                if (ty == 0) self.sphere = li; else if (ty == 1) self.plane = li;
                else if (ty == 2) self.cuboid = li; else if (ty == 3) self.tri = li; else self.aa = li;
            }
        }
        sp_intersect_chunk(s_geom, r.o, r.d, self, best);
        if (best.idx >= 0) hit_id = sp_chunk_id(s_geom, best.idx);
    }
    const uint32_t keep = __ballot_sync(0xffffffffu, active);
    if (active) {
        const uint32_t j = __popc(keep & ((1u << lane) - 1u));
        q[0 * 32 + j] = __float_as_uint(r.o.x); q[1 * 32 + j] = __float_as_uint(r.o.y); q[2 * 32 + j] = __float_as_uint(r.o.z);
        q[3 * 32 + j] = __float_as_uint(r.d.x); q[4 * 32 + j] = __float_as_uint(r.d.y); q[5 * 32 + j] = __float_as_uint(r.d.z);
        q[6 * 32 + j] = __float_as_uint(best.t);
        q[7 * 32 + j] = (hit_id < 0 ? 0x7FFFFFFFu : (uint32_t)hit_id) | (best.orient > 0 ? 0x80000000u : 0u);
        q[8 * 32 + j] = src == SP_SRC_NONE ? 0xFFFFFFFFu : src;
        q[9 * 32 + j] = mode;
        q[10 * 32 + j] = item;
    }
    __syncwarp();
    return __popc(keep);
}

template <uint32_t FEAT>
__global__ void __launch_bounds__(SPT_BLOCK, SPT_CTAS)
sp_trace_kernel(const __grid_constant__ DScene sc, const __grid_constant__ LevelArgs a) {
    extern __shared__ float4 s_geom[];
    __shared__ TraceShared sh;
    if (*reinterpret_cast<volatile const unsigned int*>(&a.out.stats->overflow) & 0xFFFFu) return;
    uint32_t n_rays = 0, fan_n[SP_MAX_FAN_CLASSES], total;
#pragma unroll
    for (int c = 0; c < SP_MAX_FAN_CLASSES; ++c) fan_n[c] = 0;
    if ((FEAT & SP_F_QUEUES) && a.source == SP_SRC_QUEUES) {
        n_rays = min(__ldg(a.in_counts), a.in_rays.capacity);
        total = n_rays;
#pragma unroll
        for (int c = 0; c < SP_MAX_FAN_CLASSES; ++c)
            if (c < sc.n_fan_classes) {
                fan_n[c] = min(__ldg(a.in_counts + 1 + c), a.in_fan_cap[c]);
                total += fan_n[c] * (uint32_t)sc.fan_mult[c];
            }
    } else {
        total = a.n_items0;
    }
    if (total == 0u || total > a.hits_cap) return;           // too many items for the hit array: the level kernel intersects itself
    for (int i = threadIdx.x, n = __ldg(sc.all.chunk_off + 1) - __ldg(sc.all.chunk_off); i < n; i += SPT_BLOCK) s_geom[i] = __ldg(sc.all.data + i);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t* const q = &sh.q[threadIdx.x >> 5][0][0];
    uint32_t* const work = const_cast<uint32_t*>(a.in_counts) + SP_COUNTS_PER_LEVEL / 2 + 1;
    const DBvh& bvh = sc.bvh;

    uint32_t q_n = 0;                                        // prepared rays waiting in the warp's queue (warp-uniform)
    uint32_t wb = 0, wend = 0;                               // the warp's current batch of items
    bool drained = false;                                    // the launch has no more items for this warp
    // the lane's ray and where its traversal stands
    bool have = false;
    float3 O = v3(0.f), D = v3(0.f), inv = v3(0.f), noi = v3(0.f);
    ChunkBest best; best.t = SP_INF; best.idx = -1; best.orient = 0;
    int src_id = -1, node = SP_BVH_DONE, sp = 0;
    uint32_t mode = 0u, item = 0u;
    int stack[32];

    while (true) {
        // ---- idle lanes pop prepared rays; the queue is refilled with full lanes when it runs dry -----------------------
        uint32_t idle = __ballot_sync(0xffffffffu, !have);
        while (idle && !(drained && q_n == 0u)) {
            if (q_n == 0u) {
                if (wb == wend) {
                    uint32_t nb = 0;
                    if (lane == 0) nb = atomicAdd(work, 32u * SPT_BATCH);
                    nb = __shfl_sync(0xffffffffu, nb, 0);
                    if (nb >= total) { drained = true; break; }
                    wb = nb; wend = min(nb + 32u * SPT_BATCH, total);
                }
                q_n = sp_trace_prepare<FEAT>(&sc, &a, s_geom, q, wb, total, n_rays, fan_n);
                wb = min(wb + 32u, wend);
                continue;
            }
            const uint32_t take = min((uint32_t)__popc(idle), q_n), rank = __popc(idle & ((1u << lane) - 1u));
            if (!have && rank < take) {
                const uint32_t j = q_n - 1u - rank;
                O = v3(__uint_as_float(q[0 * 32 + j]), __uint_as_float(q[1 * 32 + j]), __uint_as_float(q[2 * 32 + j]));
                D = v3(__uint_as_float(q[3 * 32 + j]), __uint_as_float(q[4 * 32 + j]), __uint_as_float(q[5 * 32 + j]));
                best.t = __uint_as_float(q[6 * 32 + j]);
                const uint32_t code = q[7 * 32 + j];
                best.idx = (code & 0x7FFFFFFFu) == 0x7FFFFFFFu ? -1 : (int)(code & 0x7FFFFFFFu);
                best.orient = (code & 0x80000000u) ? 1 : -1;
                src_id = (int)q[8 * 32 + j];
                mode = q[9 * 32 + j];
                item = q[10 * 32 + j];
                inv = v3(sp_safe_rcp(D.x), sp_safe_rcp(D.y), sp_safe_rcp(D.z));
                noi = v3(-O.x * inv.x, -O.y * inv.y, -O.z * inv.z);
                node = 0; sp = 0;
                have = true;
            }
            __syncwarp();
            q_n -= take;
            idle = __ballot_sync(0xffffffffu, !have);
        }
        if (__ballot_sync(0xffffffffu, have) == 0u) break;   // nothing in flight, nothing left to fetch

        // ---- one round of the traversal (sp_bvh_nearest, resumable) -----------------------------------------------------
        if (have) {
            while (node >= 0 && node != SP_BVH_DONE) {       // box nodes until a leaf (or nothing) is at hand
                const float4 n0 = __ldg(bvh.nodes + 4 * node), n1 = __ldg(bvh.nodes + 4 * node + 1);
                const float4 n2 = __ldg(bvh.nodes + 4 * node + 2), n3 = __ldg(bvh.nodes + 4 * node + 3);
                float ta, tb;
                const bool ha = sp_box_hit(v3(n0.x, n0.y, n0.z), v3(n0.w, n1.x, n1.y), inv, noi, best.t, ta);
                const bool hb = sp_box_hit(v3(n1.z, n1.w, n2.x), v3(n2.y, n2.z, n2.w), inv, noi, best.t, tb);
                int ca = __float_as_int(n3.x), cb = __float_as_int(n3.y);
                if (ha && hb) {
                    if (tb < ta) { const int c = ca; ca = cb; cb = c; }
                    stack[sp++] = cb;
                    node = ca;
                } else if (ha || hb) {
                    node = ha ? ca : cb;
                } else {
                    node = sp ? stack[--sp] : SP_BVH_DONE;
                }
            }
            if (node != SP_BVH_DONE) {                       // the leaf's colliders
                const int code = ~node, count = (code & 7) + 1;
                const float4* rec = bvh.data + (code >> 3);
                for (int i = 0; i < count; ++i) {
                    const float4 head = __ldg(rec);
                    const int kind = __float_as_int(head.x), id = __float_as_int(head.y);
                    const float4* d = rec + 1;
                    rec = d + __float_as_int(head.z);
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
                node = sp ? stack[--sp] : SP_BVH_DONE;
            }
            if (node == SP_BVH_DONE) {                       // this ray is done: its hit for the level kernel
                const uint32_t code = (best.idx < 0 ? 0x7FFFFFFFu : (uint32_t)best.idx) | (best.orient > 0 ? 0x80000000u : 0u);
                a.hits[item] = make_float2(best.t, __uint_as_float(code));
                have = false;
            }
        }
    }
}

// ---- shadow rays of Glossy hits, after the level launch (scenes behind a BVH) -----------------------------------------
// The requests the shading phase queued (LevelArgs::shq) are answered with the same persistent-lane scheme as
// sp_trace_kernel: 32 requests at a time are loaded and walked over the shadow-caster stream with full lanes, then
// every lane runs the any-hit BVH traversal of one request and takes the next when it is done; a light that is
// visible adds the radiance the request carries to its pixel (glossy.py:53-57: nearest shadow-caster distance >= the
// distance to the light).
__global__ void __launch_bounds__(SPT_BLOCK, SPT_CTAS)
sp_shadow_kernel(const __grid_constant__ DScene sc, const __grid_constant__ LevelArgs a) {
    __shared__ TraceShared sh;
    const uint32_t total = min(a.shq_count[0], a.shq_cap);
    if (total == 0u) return;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t* const q = &sh.q[threadIdx.x >> 5][0][0];
    uint32_t* const work = a.shq_count + 1;
    const DBvh& bvh = sc.bvh;
    uint32_t q_n = 0, wb = 0, wend = 0;
    bool drained = false, have = false;
    float3 O = v3(0.f), D = v3(0.f), inv = v3(0.f), noi = v3(0.f), contrib = v3(0.f);
    float dist = 0.f;
    ChunkBest best; best.t = SP_INF; best.idx = -1; best.orient = 0;
    int src_id = -1, node = SP_BVH_DONE, sp = 0;
    uint32_t mode = 0u, pix = 0u;
    int stack[32];

    while (true) {
        uint32_t idle = __ballot_sync(0xffffffffu, !have);
        while (idle && !(drained && q_n == 0u)) {
            if (q_n == 0u) {
                if (wb == wend) {
                    uint32_t nb = 0;
                    if (lane == 0) nb = atomicAdd(work, 32u * SPT_BATCH);
                    nb = __shfl_sync(0xffffffffu, nb, 0);
                    if (nb >= total) { drained = true; break; }
                    wb = nb; wend = min(nb + 32u * SPT_BATCH, total);
                }
                // prepare: load 32 requests, walk the shadow-caster stream (colliders outside the BVH) with full lanes,
                // park the ones whose light is not yet known to be hidden
                const uint32_t i = wb + lane;
                bool keep_it = false;
                float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0;
                float t_stream = SP_INF;
                if (i < wend) {
                    r0 = a.shq[3 * (size_t)i]; r1 = a.shq[3 * (size_t)i + 1]; r2 = a.shq[3 * (size_t)i + 2];
                    const uint32_t code = __float_as_uint(r2.w);
                    const int sid = (int)(code >> 2);
                    ChunkBest b0; b0.t = SP_INF; b0.idx = -1; b0.orient = 0;
                    const int2 where = a.shadow_slot ? __ldg(a.shadow_slot + sid) : make_int2(-1, -1);
                    for (int c = 0; c < sc.shadow.n_chunks; ++c) {
                        const float4* ch = sc.shadow.data + __ldg(sc.shadow.chunk_off + c);
                        SelfSlot self; self.sphere = self.plane = self.cuboid = self.tri = self.aa = -1; self.mode = code & 3u;
                        if (where.x == c) {
                            const int ty = where.y >> 28, li = where.y & 0x0FFFFFFF;
                            if (ty == 0) self.sphere = li; else if (ty == 1) self.plane = li;
                            else if (ty == 2) self.cuboid = li; else if (ty == 3) self.tri = li; else self.aa = li;
                        }
                        sp_intersect_chunk(ch, xyz(r0), xyz(r1), self, b0);
                    }
                    t_stream = b0.t;
                    keep_it = !(t_stream < r0.w);                 // already hidden by a collider of the stream: nothing to add
                }
                const uint32_t keep = __ballot_sync(0xffffffffu, keep_it);
                if (keep_it) {
                    const uint32_t j = __popc(keep & ((1u << lane) - 1u));
                    q[0 * 32 + j] = __float_as_uint(r0.x); q[1 * 32 + j] = __float_as_uint(r0.y); q[2 * 32 + j] = __float_as_uint(r0.z);
                    q[3 * 32 + j] = __float_as_uint(r1.x); q[4 * 32 + j] = __float_as_uint(r1.y); q[5 * 32 + j] = __float_as_uint(r1.z);
                    q[6 * 32 + j] = __float_as_uint(r0.w); q[7 * 32 + j] = __float_as_uint(r1.w);
                    q[8 * 32 + j] = __float_as_uint(r2.x); q[9 * 32 + j] = __float_as_uint(r2.y); q[10 * 32 + j] = __float_as_uint(r2.z);
                    q[11 * 32 + j] = __float_as_uint(r2.w);
                }
                __syncwarp();
                q_n = __popc(keep);
                wb = min(wb + 32u, wend);
                continue;
            }
            const uint32_t take = min((uint32_t)__popc(idle), q_n), rank = __popc(idle & ((1u << lane) - 1u));
            if (!have && rank < take) {
                const uint32_t j = q_n - 1u - rank;
                O = v3(__uint_as_float(q[0 * 32 + j]), __uint_as_float(q[1 * 32 + j]), __uint_as_float(q[2 * 32 + j]));
                D = v3(__uint_as_float(q[3 * 32 + j]), __uint_as_float(q[4 * 32 + j]), __uint_as_float(q[5 * 32 + j]));
                dist = __uint_as_float(q[6 * 32 + j]); pix = q[7 * 32 + j];
                contrib = v3(__uint_as_float(q[8 * 32 + j]), __uint_as_float(q[9 * 32 + j]), __uint_as_float(q[10 * 32 + j]));
                const uint32_t code = q[11 * 32 + j];
                src_id = (int)(code >> 2); mode = code & 3u;
                inv = v3(sp_safe_rcp(D.x), sp_safe_rcp(D.y), sp_safe_rcp(D.z));
                noi = v3(-O.x * inv.x, -O.y * inv.y, -O.z * inv.z);
                best.t = dist; best.idx = -1; best.orient = 0;       // only casters nearer than the light matter
                node = 0; sp = 0;
                have = true;
            }
            __syncwarp();
            q_n -= take;
            idle = __ballot_sync(0xffffffffu, !have);
        }
        if (__ballot_sync(0xffffffffu, have) == 0u) break;

        if (have) {                                              // one round of the any-hit traversal (sp_bvh_nearest, casters only)
            while (node >= 0 && node != SP_BVH_DONE) {
                const float4 n0 = __ldg(bvh.nodes + 4 * node), n1 = __ldg(bvh.nodes + 4 * node + 1);
                const float4 n2 = __ldg(bvh.nodes + 4 * node + 2), n3 = __ldg(bvh.nodes + 4 * node + 3);
                float ta, tb;
                const bool ha = sp_box_hit(v3(n0.x, n0.y, n0.z), v3(n0.w, n1.x, n1.y), inv, noi, best.t, ta);
                const bool hb = sp_box_hit(v3(n1.z, n1.w, n2.x), v3(n2.y, n2.z, n2.w), inv, noi, best.t, tb);
                int ca = __float_as_int(n3.x), cb = __float_as_int(n3.y);
                if (ha && hb) {
                    if (tb < ta) { const int c = ca; ca = cb; cb = c; }
                    stack[sp++] = cb;
                    node = ca;
                } else if (ha || hb) {
                    node = ha ? ca : cb;
                } else {
                    node = sp ? stack[--sp] : SP_BVH_DONE;
                }
            }
            if (node != SP_BVH_DONE) {
                const int code = ~node, count = (code & 7) + 1;
                const float4* rec = bvh.data + (code >> 3);
                for (int i = 0; i < count; ++i) {
                    const float4 head = __ldg(rec);
                    const int kind = __float_as_int(head.x), id = __float_as_int(head.y);
                    const float4* d = rec + 1;
                    rec = d + __float_as_int(head.z);
                    if (!(kind & 256)) continue;                     // not a shadow caster
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
                node = (best.t < dist) ? SP_BVH_DONE : (sp ? stack[--sp] : SP_BVH_DONE);     // hidden: stop at once
            }
            if (node == SP_BVH_DONE) {
                if (!(best.t < dist)) {                              // the light is visible from here
                    sp_accum_add(a.accum + pix, contrib);
                }
                have = false;
            }
        }
    }
}

#include "sp_warp_kernel.cuh"
#include "sp_split_kernels.cuh"

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

// A finished chunk's radiance moves from the scratch frame to the accumulation buffer — unless one of the chunk's queues
// overflowed (stats->overflow, set by the level kernels): what the chunk added is then incomplete, the scratch frame is
// cleared instead, and the host renders the chunk again in smaller pieces (sp_api.cu).  Deciding this on the device
// lets the host enqueue the next chunk without waiting for this one's counters.
__global__ void __launch_bounds__(256) sp_fold_kernel(float4* __restrict__ accum, float4* __restrict__ scratch, uint32_t n,
                                                      const DeviceStats* __restrict__ stats) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= n) return;
    const float4 s = scratch[i];
    if (s.x == 0.f && s.y == 0.f && s.z == 0.f) return;
    if (!(stats->overflow & 0xFFFFu)) {
        float4 a = accum[i];
        a.x += s.x; a.y += s.y; a.z += s.z;
        accum[i] = a;
    }
    scratch[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// Frame of another GPU added to this one's (sp_render_group): `other` is read straight from the peer's memory over
// NVLink when peer access is enabled — the gather and the sum are one pass, no staging copy.
__global__ void __launch_bounds__(256) sp_add_kernel(float4* __restrict__ accum, const float4* __restrict__ other, uint32_t n) {
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < n; i += gridDim.x * 256u) {
        const float4 o = other[i];
        float4 a = accum[i];
        a.x += o.x; a.y += o.y; a.z += o.z;
        accum[i] = a;
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
static size_t geom_smem_bytes(const DScene& sc) { return (size_t)sc.all.max_chunk_vec4 * sizeof(float4); }

// Instantiated material sets, smallest first; a scene runs the first one that covers its materials.
#define SP_SET_MC      (SP_F_DIFFUSE | SP_F_REFR)                                        /* Cornell box */
#define SP_SET_WHITTED (SP_F_TEX | SP_F_GLOSSY | SP_F_REFR | SP_F_THIN | SP_F_SKY)       /* examples 1-4 */
#define SP_SET_ALL     (SP_F_MATERIALS)
#define SP_SET_ALL_BVH (SP_F_MATERIALS | SP_F_BVH)                                        /* many colliders */
static const uint32_t kMaterialSets[] = {SP_SET_MC, SP_SET_WHITTED, SP_SET_ALL, SP_SET_ALL_BVH};

uint32_t sp_pick_material_set(uint32_t needed) {
    for (uint32_t set : kMaterialSets)
        if ((needed & ~set) == 0u) return set;
    return SP_SET_ALL_BVH;
}

typedef void (*LevelKernel)(const DScene, const LevelArgs);
static LevelKernel level_kernel(uint32_t material_set, bool level0) {
    switch (material_set) {
    case SP_SET_MC: return level0 ? sp_level_kernel<SP_SET_MC | SP_F_LEVEL0> : sp_level_kernel<SP_SET_MC | SP_F_QUEUES>;
    case SP_SET_WHITTED: return level0 ? sp_level_kernel<SP_SET_WHITTED | SP_F_LEVEL0> : sp_level_kernel<SP_SET_WHITTED | SP_F_QUEUES>;
    case SP_SET_ALL_BVH: return level0 ? sp_level_kernel<SP_SET_ALL_BVH | SP_F_LEVEL0> : sp_level_kernel<SP_SET_ALL_BVH | SP_F_QUEUES>;
    default: return level0 ? sp_level_kernel<SP_SET_ALL | SP_F_LEVEL0> : sp_level_kernel<SP_SET_ALL | SP_F_QUEUES>;
    }
}

// Queue-fed levels of small untextured Monte-Carlo scenes run the warp-autonomous kernel (sp_warp_kernel.cuh).
// SIGHTPY_WARP_KERNEL=0 / option "warp_kernel" = 0 keeps them on sp_level_kernel (A/B measurements, parity tests).
bool sp_use_warp_kernel(const DScene& sc, uint32_t material_set) {
    static const bool enabled = [] { const char* e = getenv("SIGHTPY_WARP_KERNEL"); return !(e && e[0] == '0'); }();
    if (!enabled || !sc.use_warp_kernel || material_set != SP_SET_MC) return false;
    if (sc.all.n_chunks != 1 || sc.bvh.n_nodes != 0 || sc.n_colliders > SPW_MAX_COLLIDERS) return false;
    for (int c = 0; c < sc.n_fan_classes; ++c)
        if (sc.fan_mult[c] > 1024) return false;
    return true;
}

int sp_level_grid(int device, const DScene& sc, uint32_t material_set, bool level0) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (!level0 && sp_use_warp_kernel(sc, material_set)) {
        auto k = sp_warp_kernel<SP_SET_MC>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SP_CHUNK_VEC4 * sizeof(float4)));
        int per_sm = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, SPW_BLOCK, geom_smem_bytes(sc)) != cudaSuccess || per_sm < 1)
            per_sm = 1;
        return sms * per_sm;
    }
    LevelKernel k = level_kernel(material_set, level0);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SP_CHUNK_VEC4 * sizeof(float4)));
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, SP_BLOCK, geom_smem_bytes(sc)) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    return sms * per_sm;
}

cudaError_t sp_launch_level(const DScene& sc, const LevelArgs& a, uint32_t material_set, int grid, cudaStream_t st) {
    if (a.source == SP_SRC_QUEUES && a.run == SP_RUN_FULL && sp_use_warp_kernel(sc, material_set)) {
        sp_warp_kernel<SP_SET_MC><<<grid, SPW_BLOCK, geom_smem_bytes(sc), st>>>(sc, a);
        return cudaGetLastError();
    }
    level_kernel(material_set, a.source != SP_SRC_QUEUES)<<<grid, SP_BLOCK, geom_smem_bytes(sc), st>>>(sc, a);
    return cudaGetLastError();
}

bool sp_can_pretrace(const DScene& sc, uint32_t material_set) {
    static const bool enabled = [] { const char* e = getenv("SIGHTPY_PRETRACE"); return !(e && e[0] == '0'); }();
    return enabled && material_set == SP_SET_ALL_BVH && sc.bvh.n_nodes > 0 && sc.all.n_chunks == 1;
}

cudaError_t sp_launch_shadow(const DScene& sc, const LevelArgs& a, int device, cudaStream_t st) {
    static int grid[16] = {0};
    if (grid[device & 15] == 0) {
        int sms = 148, per_sm = 1;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sp_shadow_kernel, SPT_BLOCK, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
        grid[device & 15] = sms * per_sm;
    }
    sp_shadow_kernel<<<grid[device & 15], SPT_BLOCK, 0, st>>>(sc, a);
    return cudaGetLastError();
}

cudaError_t sp_launch_trace(const DScene& sc, const LevelArgs& a, uint32_t material_set, int device, cudaStream_t st) {
    (void)material_set;
    static int grid[16] = {0};
    auto k0 = sp_trace_kernel<SP_SET_ALL_BVH | SP_F_LEVEL0>;
    auto kq = sp_trace_kernel<SP_SET_ALL_BVH | SP_F_QUEUES>;
    auto k = a.source != SP_SRC_QUEUES ? k0 : kq;
    if (grid[device & 15] == 0) {
        int sms = 148, per_sm = 1;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SP_CHUNK_VEC4 * sizeof(float4)));
        cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SP_CHUNK_VEC4 * sizeof(float4)));
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kq, SPT_BLOCK, (size_t)SP_CHUNK_VEC4 * sizeof(float4)) != cudaSuccess || per_sm < 1) per_sm = 1;
        grid[device & 15] = sms * per_sm;
    }
    k<<<grid[device & 15], SPT_BLOCK, geom_smem_bytes(sc), st>>>(sc, a);
    return cudaGetLastError();
}

// Whitted scenes (no Diffuse fans, no BVH, one staged chunk) run a level as hit kernel + per-material shade kernels.
// SIGHTPY_SPLIT=0 / option "split_kernels" = 0 keeps them on the fused sp_level_kernel (A/B measurements, parity tests).
bool sp_can_split(const DScene& sc, uint32_t material_set) {
    static const bool enabled = [] { const char* e = getenv("SIGHTPY_SPLIT"); return !(e && e[0] == '0'); }();
    return enabled && sc.use_split && material_set == SP_SET_WHITTED && sc.all.n_chunks == 1 && sc.bvh.n_nodes == 0;
}

template <uint32_t SRC>
static cudaError_t launch_split(const DScene& sc, const LevelArgs& a, uint32_t kind_mask, int grid, cudaStream_t st, int* launched) {
    auto kh = sp_hit_kernel<SP_SET_WHITTED | SRC>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(sp_hit_kernel<SP_SET_WHITTED | SP_F_LEVEL0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SP_CHUNK_VEC4 * sizeof(float4)));
        cudaFuncSetAttribute(sp_hit_kernel<SP_SET_WHITTED | SP_F_QUEUES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SP_CHUNK_VEC4 * sizeof(float4)));
        attr_set = true;
    }
    int n = 1;
    kh<<<grid, SPS_BLOCK, geom_smem_bytes(sc), st>>>(sc, a);
    if (kind_mask & (1u << SP_MAT_REFRACTIVE)) { sp_shade_kernel<SP_F_TEX | SP_F_REFR | SRC, 0><<<grid, SPS_BLOCK, 0, st>>>(sc, a); ++n; }
    if (kind_mask & (1u << SP_MAT_GLOSSY)) { sp_shade_kernel<SP_F_TEX | SP_F_GLOSSY | SRC, 1><<<grid, SPS_BLOCK, 0, st>>>(sc, a); ++n; }
    if (kind_mask & (1u << SP_MAT_THINFILM)) { sp_shade_kernel<SP_F_TEX | SP_F_THIN | SRC, 2><<<grid, SPS_BLOCK, 0, st>>>(sc, a); ++n; }
    if (kind_mask & (1u << SP_MAT_SKYBOX)) { sp_shade_kernel<SP_F_TEX | SP_F_SKY | SRC, 4><<<grid, SPS_BLOCK, 0, st>>>(sc, a); ++n; }
    if (kind_mask & (1u << SP_MAT_EMISSIVE)) { sp_shade_kernel<SP_F_TEX | SRC, 5><<<grid, SPS_BLOCK, 0, st>>>(sc, a); ++n; }
    if (launched) *launched = n;
    return cudaGetLastError();
}

cudaError_t sp_launch_split_level(const DScene& sc, const LevelArgs& a, uint32_t kind_mask, int device, cudaStream_t st, int* launched) {
    static int sms[16] = {0};
    if (sms[device & 15] == 0) {
        int v = 148;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
        sms[device & 15] = v;
    }
    const int grid = sms[device & 15] * 8;              // 4 resident CTAs per SM, two waves: grid-stride over the items
    return a.source != SP_SRC_QUEUES ? launch_split<SP_F_LEVEL0>(sc, a, kind_mask, grid, st, launched)
                                     : launch_split<SP_F_QUEUES>(sc, a, kind_mask, grid, st, launched);
}

cudaError_t sp_launch_resolve(const ResolveArgs& a, cudaStream_t st) {
    if (a.n_pix == 0) return cudaSuccess;
    sp_resolve_kernel<<<(a.n_pix + 255u) / 256u, 256, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t sp_launch_fold(float4* accum, float4* scratch, uint32_t n_pix, const DeviceStats* stats, cudaStream_t st) {
    if (n_pix == 0) return cudaSuccess;
    sp_fold_kernel<<<(n_pix + 255u) / 256u, 256, 0, st>>>(accum, scratch, n_pix, stats);
    return cudaGetLastError();
}

cudaError_t sp_launch_add(float4* accum, const float4* other, uint32_t n_pix, cudaStream_t st) {
    if (n_pix == 0) return cudaSuccess;
    sp_add_kernel<<<148 * 8, 256, 0, st>>>(accum, other, n_pix);
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
