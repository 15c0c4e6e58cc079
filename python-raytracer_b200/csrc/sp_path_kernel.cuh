// Path kernel: the warp-autonomous, persistent-lane form of the wavefront level (round 2).
//
// Same contract as sp_level_kernel (sp_kernels.cu): one launch consumes the work of one wavefront level of
// get_raycolor (ray.py:122-148) — camera / caller rays at level 0, the previous launch's queue records
// afterwards — and appends the records of the next level.  What differs from round 1's kernels:
//
//   * Every lane owns a *path*, not an item of a 32-wide batch.  A lane whose path ended picks up new work at the
//     top of the next iteration ("regeneration"), so the lanes of a warp stay full no matter how the paths
//     of its rays differ in length.
//   * A Diffuse hit that continues with a single ray (diffuse.py:85-121, the second bounce of the Cornell
//     box: two thirds of that frame's rays) is *not* written to a queue: hit point, shading normal and
//     throughput stay in the lane's registers and the loop goes back to "sample a direction, intersect".
//     One copy of the sampling / intersection code serves fresh and continued rays alike.
//   * A Diffuse hit that fans out (diffuse.py:34-83, diffuse_rays = 20 children) becomes the lane's *parent*:
//     the record lives in the lane's shared-memory slot and the lane walks its children one after the other.
//     Level-0 launches adopt the fan of their own primary hit, so camera ray -> 20 children -> 20 second bounces
//     run in one launch without touching HBM; queue-fed launches read parents the previous launch queued
//     (glass -> diffuse paths), prefetched into the slot with cp.async while the lane still traces.
//   * Everything else a hit can be (Glossy with its shadow rays, Refractive, ThinFilm, SkyBox, textured
//     Diffuse / Emissive) goes into a *warp-private stash per material kind* and is shaded 32 hits at a time
//     by a non-inlined sp_shade instantiation for just that kind, so the expensive code runs with full lanes
//     and only the kinds a scene uses occupy shared memory or the instruction cache.  Their children go to the
//     next level's queue through warp-private slabs (one global atomic per 32-256 records).
//   * Work is drawn dynamically: a warp takes batches of items from a per-segment counter and hands them to
//     its lanes as they fall idle.
// Rays, random numbers (Philox counters: pixel, path-tree node, block) and results are those of
// sp_level_kernel; option "warp_kernel" = 0 keeps a scene on that kernel (tests compare the two).
#pragma once
#include "sp_launch.h"
#include "sp_sampling.cuh"
#include "sp_shade.cuh"

#ifndef SPP_BLOCK
#define SPP_BLOCK 256
#endif
#define SPP_WARPS (SPP_BLOCK / 32)
#ifndef SPP_CTAS_MC
#define SPP_CTAS_MC 4                // untextured Diffuse / Refractive / Emissive: 64 registers, one stash
#endif
#ifndef SPP_CTAS_FULL
#define SPP_CTAS_FULL 2              // textured / glossy sets: up to 128 registers, shared memory for several stashes
#endif
#define SPP_CTAS(FEAT) ((((FEAT) & (SP_F_TEX | SP_F_GLOSSY | SP_F_THIN | SP_F_SKY | SP_F_BVH)) == 0u) ? SPP_CTAS_MC : SPP_CTAS_FULL)
#define SPP_STASH_WORDS 14           // o d thr pix path meta t (id | outer << 31)
#define SPP_STASH_CAP 64             // < 32 left over + 32 pushed
#define SPP_STASH_STRIDE (SPP_WARPS * SPP_STASH_WORDS * SPP_STASH_CAP)     // words per bin (all warps of the CTA)
#define SPP_N_QUEUES (1 + SP_MAX_FAN_CLASSES)
#define SPP_ITEM_NONE 0xFFFFFFFFu

// Development build (-DSP_CHECKED): the hand-rolled protocols assert their invariants and report through the
// overflow word (bits 16+), which the host turns into an error.
#ifdef SP_CHECKED
#define SP_ASSERT(stats, cond, code) do { if (!(cond)) atomicOr(&(stats)->overflow, 0x10000u << (code)); } while (0)
#else
#define SP_ASSERT(stats, cond, code) do { } while (0)
#endif
enum { SP_CHK_SLOT = 0, SP_CHK_STASH = 1, SP_CHK_SLAB = 2, SP_CHK_CHILD = 3, SP_CHK_DEPTH = 4 };

struct PathShared {
    float4 rec[SPP_WARPS][3][32];                  // per lane: its parent record (fan) / next queue record (cp.async target)
    uint32_t slab[SPP_WARPS][SPP_N_QUEUES][2];     // per warp and output queue: next free slot, end of the slab
    uint32_t nst[SPP_WARPS][8];                    // per warp and stash bin: entries waiting
    uint32_t hist[SPP_WARPS][SP_MAX_LEVELS];       // per warp: rays traced per ray depth
    uint32_t seg[SPP_N_QUEUES][8];                 // per work segment: SPP_SEG_* constants
    uint2 cls[SP_SMALL_COLLIDERS];                 // per collider: what a hit does (DScene::col_cls)
    float2 src_info[SP_SMALL_COLLIDERS];           // per collider: position in the chunk's id array (int bits), cosine-pdf weight
    float4 lite[SP_SMALL_COLLIDERS];               // per collider: albedo / emitted colour, 1 / diffuse_rays
    float4 imp[SP_MAX_IMPORTANCE];                 // importance list (centre, radius)
};
enum { SPP_SEG_ITEMS = 0, SPP_SEG_MULT, SPP_SEG_BASE, SPP_SEG_SLAB, SPP_SEG_BATCH };

SP_DEV void spp_cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
SP_DEV void spp_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
SP_DEV void spp_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
SP_DEV uint32_t spp_lane_id() { uint32_t r; asm("mov.u32 %0, %%laneid;" : "=r"(r)); return r; }
SP_DEV uint32_t spp_lanemask_lt() { uint32_t r; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(r)); return r; }

// ---- queue slots from warp-private slabs ------------------------------------------------------------------
// Slots for `tot` records of output queue q, requested by the whole warp at once.  Rank x of the request
// lives at  x < rem ? first + x : fresh + (x - rem).
struct PathGrant { uint32_t first, rem, fresh; };
SP_DEV uint32_t spp_slab_pos(const PathGrant& g, uint32_t x) {
    if (x < g.rem) return g.first + x;
    return g.fresh == SP_SLOT_NONE ? SP_SLOT_NONE : g.fresh + (x - g.rem);
}

SP_DEV PathGrant spp_slab_alloc(uint32_t* slab, uint32_t tot, uint32_t q, uint32_t slab_size, const LevelOut& out, uint32_t lane) {
    PathGrant g;
    uint2 st = *reinterpret_cast<const uint2*>(slab);         // x = next free slot, y = end of the slab
    SP_ASSERT(out.stats, st.x <= st.y, SP_CHK_SLAB);
    g.first = st.x; g.rem = st.y - st.x; g.fresh = SP_SLOT_NONE;
    if (tot > g.rem) {                                        // warp-uniform: finish this slab, open another
        uint32_t b = 0;
        if (lane == 0) {
            const uint32_t cap = (q == 0) ? out.rays.capacity : out.fan_cap[q - 1];
            b = atomicAdd(out.counts + q, slab_size);
            if (b + slab_size > cap || b + slab_size < b) { atomicOr(&out.stats->overflow, 1u); b = SP_SLOT_NONE; }
            else if (q > 0) b += out.fan_base[q - 1];
        }
        b = __shfl_sync(0xffffffffu, b, 0);
        g.fresh = b;
        st.x = b + (tot - g.rem); st.y = b + slab_size;
        if (b == SP_SLOT_NONE) st.x = st.y = 0u;
    } else {
        st.x += tot;
    }
    __syncwarp();
    *reinterpret_cast<uint2*>(slab) = st;                     // every lane writes the same value
    __syncwarp();
    return g;
}

// ---- one stash bin, 32 hits at a time --------------------------------------------------------------------------
// Shade the entries [first, first + n) (n <= 32) of one warp-private stash with sp_shade<SFEAT>, SFEAT being the
// feature subset of that bin's material kind.
template <uint32_t SFEAT>
__device__ __noinline__ void spp_shade_stash(const DScene* scp, const LevelArgs* ap, uint32_t* stash, uint32_t* slabs,
                                             const float* lin_lut, uint32_t first, uint32_t n, uint32_t slab_size) {
    const DScene& sc = *scp;
    const LevelArgs& a = *ap;
    const uint32_t lane = spp_lane_id();
    const bool mine = lane < n;
    Ray s;
    HitRec h;
    s.o = s.d = s.thr = v3(0.f); s.pix = s.path = s.meta = 0u; h.t = 0.f; h.id = 0; h.orient = 1;
    int n_ray = 0, fan_class = -1;
    if (mine) {
        const uint32_t j = first + lane;
        SP_ASSERT(a.out.stats, j < SPP_STASH_CAP, SP_CHK_STASH);
        const uint32_t* st = stash + j;
        s.o = v3(__uint_as_float(st[0 * SPP_STASH_CAP]), __uint_as_float(st[1 * SPP_STASH_CAP]), __uint_as_float(st[2 * SPP_STASH_CAP]));
        s.d = v3(__uint_as_float(st[3 * SPP_STASH_CAP]), __uint_as_float(st[4 * SPP_STASH_CAP]), __uint_as_float(st[5 * SPP_STASH_CAP]));
        s.thr = v3(__uint_as_float(st[6 * SPP_STASH_CAP]), __uint_as_float(st[7 * SPP_STASH_CAP]), __uint_as_float(st[8 * SPP_STASH_CAP]));
        s.pix = st[9 * SPP_STASH_CAP]; s.path = st[10 * SPP_STASH_CAP]; s.meta = st[11 * SPP_STASH_CAP];
        h.t = __uint_as_float(st[12 * SPP_STASH_CAP]);
        const uint32_t packed = st[13 * SPP_STASH_CAP];
        h.id = (int)(packed & 0x7FFFFFFFu); h.orient = (packed & 0x80000000u) ? 1 : -1;
        const float4 raw = __ldg(reinterpret_cast<const float4*>(sc.col_info + h.id));
        const DColInfo ci = *reinterpret_cast<const DColInfo*>(&raw);
        sp_child_needs(ci, meta_depth(s.meta), meta_dr(s.meta), n_ray, fan_class);
    }
    __syncwarp();                                             // the entries are free again
    const uint32_t lt_mask = spp_lanemask_lt();
    ShadeCtx ctx;
    ctx.sc = scp; ctx.out = &a.out; ctx.shadow_slot = a.shadow_slot; ctx.lin_lut = lin_lut; ctx.shadow_rays = 0;
    ctx.ray_slot = ctx.ray_slot1 = SP_SLOT_NONE; ctx.ray_used = 0u; ctx.fan_slot = SP_SLOT_NONE;
    if (SFEAT & (SP_F_GLOSSY | SP_F_REFR | SP_F_THIN)) {
        const uint32_t b0 = __ballot_sync(0xffffffffu, n_ray & 1), b1 = __ballot_sync(0xffffffffu, n_ray & 2);
        const uint32_t tot = __popc(b0) + 2u * __popc(b1);
        if (tot) {
            const PathGrant g = spp_slab_alloc(slabs, tot, 0u, slab_size < 64u ? 64u : slab_size, a.out, lane);
            const uint32_t rank = __popc(b0 & lt_mask) + 2u * __popc(b1 & lt_mask);
            if (n_ray >= 1) ctx.ray_slot = spp_slab_pos(g, rank);
            if (n_ray >= 2) ctx.ray_slot1 = spp_slab_pos(g, rank + 1u);
        }
    }
    if (SFEAT & SP_F_DIFFUSE) {
        uint32_t todo = __ballot_sync(0xffffffffu, fan_class >= 0);
        while (todo) {                                        // one round per fan class present in the warp
            const int c = __shfl_sync(0xffffffffu, fan_class, __ffs(todo) - 1);
            const uint32_t bc = __ballot_sync(0xffffffffu, fan_class == c);
            todo &= ~bc;
            const PathGrant g = spp_slab_alloc(slabs + 2 * (1 + c), __popc(bc), 1u + (uint32_t)c, slab_size, a.out, lane);
            if (fan_class == c) ctx.fan_slot = spp_slab_pos(g, __popc(bc & lt_mask));
        }
    }
    const uint32_t fan_reserved = ctx.fan_slot;
    if (mine) {
        const float3 add = sp_shade<SFEAT>(ctx, s, h);
        float* px = reinterpret_cast<float*>(a.accum + s.pix);
        if (add.x != 0.f) atomicAdd(px, add.x);
        if (add.y != 0.f) atomicAdd(px + 1, add.y);
        if (add.z != 0.f) atomicAdd(px + 2, add.z);
        // reserved but unused slots become dead records
        if (ctx.ray_used < 1u && n_ray >= 1 && ctx.ray_slot != SP_SLOT_NONE) sp_write_dead(a.out.rays, ctx.ray_slot);
        if (ctx.ray_used < 2u && n_ray >= 2 && ctx.ray_slot1 != SP_SLOT_NONE) sp_write_dead(a.out.rays, ctx.ray_slot1);
        if ((SFEAT & SP_F_DIFFUSE) && fan_reserved != SP_SLOT_NONE && ctx.fan_slot != SP_SLOT_NONE) sp_write_dead(a.out.fans, fan_reserved);
    }
    if (SFEAT & SP_F_GLOSSY) {
        const uint32_t shr = __reduce_add_sync(0xffffffffu, (uint32_t)ctx.shadow_rays);
        if (lane == 0 && shr) atomicAdd(&a.out.stats->shadow_rays, (unsigned long long)shr);
    }
}

// feature subset sp_shade needs for the hits of one stash bin
template <uint32_t FEAT, int BIN> struct SppBinFeat {
    static constexpr uint32_t common = FEAT & (SP_F_TEX | SP_F_BVH);
    static constexpr uint32_t value =
        BIN == SP_BIN_REFR ? (SP_F_REFR | common) :
        BIN == SP_BIN_GLOSSY ? (SP_F_GLOSSY | common) :
        BIN == SP_BIN_THIN ? (SP_F_THIN | common) :
        BIN == SP_BIN_SKY ? (SP_F_SKY | common) : (SP_F_DIFFUSE | common);
    static constexpr bool present =
        BIN == SP_BIN_REFR ? (FEAT & SP_F_REFR) != 0u :
        BIN == SP_BIN_GLOSSY ? (FEAT & SP_F_GLOSSY) != 0u :
        BIN == SP_BIN_THIN ? (FEAT & SP_F_THIN) != 0u :
        BIN == SP_BIN_SKY ? (FEAT & SP_F_SKY) != 0u : (FEAT & SP_F_TEX) != 0u;     // textured Diffuse / Emissive
};

template <uint32_t FEAT>
SP_DEV void spp_flush_bin(uint32_t bin, const DScene* scp, const LevelArgs* ap, uint32_t* stash, uint32_t* slabs,
                          const float* lin_lut, uint32_t first, uint32_t n, uint32_t slab_size) {
    switch (bin) {
    case SP_BIN_REFR:
        if constexpr (SppBinFeat<FEAT, SP_BIN_REFR>::present) spp_shade_stash<SppBinFeat<FEAT, SP_BIN_REFR>::value>(scp, ap, stash, slabs, lin_lut, first, n, slab_size);
        break;
    case SP_BIN_GLOSSY:
        if constexpr (SppBinFeat<FEAT, SP_BIN_GLOSSY>::present) spp_shade_stash<SppBinFeat<FEAT, SP_BIN_GLOSSY>::value>(scp, ap, stash, slabs, lin_lut, first, n, slab_size);
        break;
    case SP_BIN_THIN:
        if constexpr (SppBinFeat<FEAT, SP_BIN_THIN>::present) spp_shade_stash<SppBinFeat<FEAT, SP_BIN_THIN>::value>(scp, ap, stash, slabs, lin_lut, first, n, slab_size);
        break;
    case SP_BIN_SKY:
        if constexpr (SppBinFeat<FEAT, SP_BIN_SKY>::present) spp_shade_stash<SppBinFeat<FEAT, SP_BIN_SKY>::value>(scp, ap, stash, slabs, lin_lut, first, n, slab_size);
        break;
    default:
        if constexpr (SppBinFeat<FEAT, SP_BIN_GENERIC>::present) spp_shade_stash<SppBinFeat<FEAT, SP_BIN_GENERIC>::value>(scp, ap, stash, slabs, lin_lut, first, n, slab_size);
        break;
    }
}

template <uint32_t FEAT>
__global__ void __launch_bounds__(SPP_BLOCK, SPP_CTAS(FEAT))
sp_path_kernel(const __grid_constant__ DScene sc, const __grid_constant__ LevelArgs a) {
    constexpr bool L0 = (FEAT & SP_F_LEVEL0) != 0u;          // camera / caller rays; otherwise queue records
    constexpr bool HAS_DIFF = (FEAT & SP_F_DIFFUSE) != 0u;
    constexpr bool BIG = (FEAT & SP_F_BVH) != 0u;            // more than SP_SMALL_COLLIDERS colliders: tables stay in global memory
    constexpr bool SLOTS = !L0 || HAS_DIFF;                  // lanes keep records in their shared-memory slot
    extern __shared__ float4 s_dyn[];                        // staged geometry chunk, then the stashes
    __shared__ PathShared sh;
    __shared__ float s_lin_lut[(FEAT & SP_F_TEX) ? 256 : 1];

    // a queue of an earlier level overflowed: its records are incomplete, the host discards the chunk
    if (*reinterpret_cast<volatile const unsigned int*>(&a.out.stats->overflow)) return;

    const uint32_t tid = threadIdx.x;

    // ---- work items of this launch ---------------------------------------------------------------------------
    uint32_t n_rays = 0, total = 0;
    if (L0) {
        total = a.n_items0;
    } else {
        n_rays = min(__ldg(a.in_counts), a.in_rays.capacity);
        total = n_rays;
        for (int c = 0; c < sc.n_fan_classes; ++c) total += min(__ldg(a.in_counts + 1 + c), a.in_fan_cap[c]);
    }
    if (total == 0u) return;

    sp_stage_chunk(s_dyn, sc, sc.all, 0);
    if (!BIG) {
        for (uint32_t i = tid; i < (uint32_t)sc.n_colliders; i += SPP_BLOCK) {
            sh.cls[i] = __ldg(sc.col_cls + i);
            sh.src_info[i] = __ldg(sc.col_src + i);
            sh.lite[i] = __ldg(sc.col_lite + i);
        }
    }
    if (tid < (uint32_t)sc.n_importance)
        sh.imp[tid] = make_float4(sc.importance[tid].center.x, sc.importance[tid].center.y, sc.importance[tid].center.z, sc.importance[tid].radius);
    if (FEAT & SP_F_TEX)
        for (uint32_t i = tid; i < 256u; i += SPP_BLOCK) s_lin_lut[i] = c_decode[SP_DECODE_LINEAR][i];
    for (uint32_t i = tid; i < SPP_WARPS * SPP_N_QUEUES * 2; i += SPP_BLOCK) reinterpret_cast<uint32_t*>(sh.slab)[i] = 0u;
    for (uint32_t i = tid; i < SPP_WARPS * 8; i += SPP_BLOCK) reinterpret_cast<uint32_t*>(sh.nst)[i] = 0u;
    for (uint32_t i = tid; i < SPP_WARPS * SP_MAX_LEVELS; i += SPP_BLOCK) reinterpret_cast<uint32_t*>(sh.hist)[i] = 0u;
    const int n_seg = L0 ? 1 : 1 + sc.n_fan_classes;
    if (tid < (uint32_t)n_seg) {
        // segment 0: camera / caller rays (level 0) or explicit ray records; segment 1 + c: fan parents of class c
        uint32_t mult = 1u, n_items = L0 ? a.n_items0 : n_rays, base = 0u;
        if (tid > 0) {
            mult = (uint32_t)sc.fan_mult[tid - 1];
            n_items = min(__ldg(a.in_counts + tid), a.in_fan_cap[tid - 1]);
            base = a.in_fan_base[tid - 1];
        }
        // rays an item stands for, roughly: a fan parent is `mult` children plus their second bounces; a primary
        // of a scene with fans adopts one
        uint32_t weight = mult > 1u ? 2u * mult : 1u;
        if (L0 && HAS_DIFF && sc.n_fan_classes > 1) weight = 2u * (uint32_t)sc.fan_mult[1];
        // slots per slab: about an eighth of what a warp can emit in this launch (bounded by the rays it traces), so
        // that the unused tails stay a few per cent of the queue even for small launches; a power of two in [32, 256]
        uint32_t slab_size = 32u;
        const unsigned long long per_warp = (unsigned long long)total * weight / (gridDim.x * SPP_WARPS * 8u);
        while (slab_size < 256u && slab_size * 2u <= per_warp) slab_size *= 2u;
        uint32_t* sg = sh.seg[tid];
        sg[SPP_SEG_ITEMS] = n_items; sg[SPP_SEG_MULT] = mult; sg[SPP_SEG_BASE] = base; sg[SPP_SEG_SLAB] = slab_size;
        // items per draw from the segment's counter: about 256 rays' worth (one same-address atomic per 256 rays was
        // measured to be free, one per 128 cost 4 %)
        sg[SPP_SEG_BATCH] = weight >= 8u ? 32u : 256u;
    }
    __syncthreads();

    // a warp-wide reduction hands the warp's index to the uniform datapath: everything derived from it (slot, stash,
    // slab addresses) then lives in uniform registers instead of one copy per lane
    const uint32_t warp = __reduce_max_sync(0xffffffffu, tid >> 5);
    const uint32_t lane = spp_lane_id();
    float4* const my_rec = &sh.rec[warp][0][lane];
    uint32_t* const slabs = &sh.slab[warp][0][0];
    uint32_t* const nst = sh.nst[warp];
    uint32_t* const hist = sh.hist[warp];
    uint32_t* const stash0 = reinterpret_cast<uint32_t*>(s_dyn + sc.all.max_chunk_vec4) + warp * (SPP_STASH_WORDS * SPP_STASH_CAP);
    const int* const chunk_ids = reinterpret_cast<const int*>(s_dyn + reinterpret_cast<const GeomChunkHeader*>(s_dyn)->off_ids);
    auto imp = [&](int i) { return sh.imp[i]; };

    // ---- lane state -------------------------------------------------------------------------------------------
    // The record in r is either a ray (o, d = direction) or a fan (o, d = shading normal to sample around); either
    // way thr, pix, path, meta are those of the ray to trace.  flags: bit 0 = r is valid, bit 1 = d is a direction.
    Ray r;
    r.o = r.d = r.thr = v3(0.f); r.pix = 0; r.path = 0; r.meta = 0;
    uint32_t flags = 0u;
    uint32_t child = 0u, slot_mult = 0u;                      // the slot holds a record with children [child, slot_mult) left
    uint32_t next_item = SPP_ITEM_NONE;                       // level 0: the camera / caller ray this lane traces next

    for (int seg = 0; seg < n_seg; ++seg) {
        const volatile uint32_t* sg = sh.seg[seg];
        const uint32_t n_items = sg[SPP_SEG_ITEMS];
        if (n_items == 0u) continue;
        const uint32_t batch = sg[SPP_SEG_BATCH];
        // Work distribution: a warp draws batches of consecutive items from a per-segment counter (the level's
        // queue-count block holds it, zeroed by the host with the counts) and hands them to lanes as they fall idle;
        // the counter's round trip is hidden by drawing the next batch while the current one is handed out.
        uint32_t* const work = const_cast<uint32_t*>(a.in_counts) + SPP_N_QUEUES + seg;
        uint32_t pending = 0;                                  // lane 0: start of the batch after the current one
        auto draw = [&]() { if (lane == 0) pending = atomicAdd(work, batch); };
        auto drawn = [&]() { return __reduce_max_sync(0xffffffffu, lane == 0 ? pending : 0u); };
        draw();
        uint32_t wb = drawn();
        bool drained = wb >= n_items;
        uint32_t wend = drained ? wb : min(wb + batch, n_items);
        if (!drained) draw();
        child = slot_mult = 0u;

#pragma unroll 1
        while (true) {
            // ---- (a) idle lanes take the next child of their slot's record, or (level 0) their next primary ray ----
            if (SLOTS && !L0) spp_cp_async_wait_all();
            if (!(flags & 1u)) {
                if (SLOTS && child < slot_mult) {
                    const float4 q0 = my_rec[0], q1 = my_rec[32], q2 = my_rec[64];
                    const uint32_t meta = __float_as_uint(q2.w);
                    if (meta == SP_META_DEAD) {
                        child = slot_mult;                     // a dead record has no children
                    } else {
                        const bool is_fan = L0 || seg != 0;    // the slots of a level-0 launch only ever hold adopted fans
                        r.o = xyz(q0); r.d = xyz(q1); r.thr = xyz(q2);
                        r.pix = __float_as_uint(q0.w); r.meta = meta;
                        r.path = is_fan ? sp_child_path(__float_as_uint(q1.w), child) : __float_as_uint(q1.w);
                        flags = is_fan ? 1u : 3u;
                        child += 1u;
                    }
                } else if (L0 && next_item != SPP_ITEM_NONE) {
                    if (a.source == SP_SRC_CAMERA) {
                        const uint32_t sample = a.sample_begin + next_item / a.n_pix;
                        r.pix = a.pix_begin + next_item % a.n_pix;
                        r.path = sp_root_path(sample);
                        sp_camera_ray(sc.cam, r.pix, sample, sc.seed_lo, sc.seed_hi, r.o, r.d);
                    } else {
                        const size_t i = (size_t)a.user_base + next_item;
                        r.pix = (uint32_t)i;
                        r.path = sp_root_path(0u);
                        r.o = v3(__ldg(a.user_o + 3 * i), __ldg(a.user_o + 3 * i + 1), __ldg(a.user_o + 3 * i + 2));
                        r.d = v3(__ldg(a.user_d + 3 * i), __ldg(a.user_d + 3 * i + 1), __ldg(a.user_d + 3 * i + 2));
                    }
                    r.thr = v3(1.f);
                    r.meta = sp_pack_meta(0u, 0u, 0u, SP_SRC_NONE, SP_SELF_SKIP);
                    flags = 3u;
                    next_item = SPP_ITEM_NONE;
                }
            }
            // ---- (b) lanes whose slot (level 0: next item) is used up get the next item of the segment ----------------
            if (!drained) {
                const bool want = L0 ? (next_item == SPP_ITEM_NONE) : (child >= slot_mult);
                const uint32_t need = __ballot_sync(0xffffffffu, want);
                if (need) {
                    if (wb == wend) {                          // this batch is handed out: switch to the one drawn earlier
                        const uint32_t nb = drawn();
                        if (nb < n_items) { wb = nb; wend = min(nb + batch, n_items); draw(); }
                        else drained = true;
                    }
                    const uint32_t take = min((uint32_t)__popc(need), wend - wb);
                    const uint32_t rank = __popc(need & spp_lanemask_lt());
                    if (want && rank < take) {
                        const uint32_t item = wb + rank;
                        if (L0) {
                            next_item = item;
                        } else {
                            // the record the lane needs after the ray it traces now: copied into its slot in the background
                            const RayQueue* q = seg == 0 ? &a.in_rays : &a.in_fans;
                            const uint32_t rec = sg[SPP_SEG_BASE] + item;
                            spp_cp_async16(my_rec, q->q0 + rec); spp_cp_async16(my_rec + 32, q->q1 + rec); spp_cp_async16(my_rec + 64, q->q2 + rec);
                            child = 0u; slot_mult = sg[SPP_SEG_MULT];
                        }
                    }
                    wb += take;
                }
            }
            if (SLOTS && !L0) spp_cp_async_commit();
            if (__ballot_sync(0xffffffffu, flags & 1u) == 0u) {
                const bool more = (SLOTS && child < slot_mult) || (L0 && next_item != SPP_ITEM_NONE);
                if (drained && !__any_sync(0xffffffffu, more)) break;
                continue;
            }

            // ---- 1. the ray: a fan samples its direction (diffuse.py:49-61), a ray record brings it along --------------
            int self_tag = -1;                                 // the source collider's position in the chunk's id array
            if (HAS_DIFF && (flags & 3u) == 1u) {
                const float2 si = BIG ? __ldg(sc.col_src + meta_src(r.meta)) : sh.src_info[meta_src(r.meta)];   // fans always name their source
                self_tag = __float_as_int(si.x);
                float3 dir;
                const float weight = sp_sample_diffuse_with(sc, imp, r.o, r.d, si.y, r.pix, r.path, dir);
                r.d = dir;
                r.thr = r.thr * weight;
                if (!(weight > 0.f)) flags = 0u;               // zero-weight samples cannot contribute: not traced
            } else if (flags & 1u) {
                const uint32_t src = meta_src(r.meta);
                if (src != SP_SRC_NONE) self_tag = __float_as_int((BIG ? __ldg(sc.col_src + src) : sh.src_info[src]).x);
            }
            const bool active = (flags & 1u) != 0u;
            flags = 0u;                                        // consumed; a continuing Diffuse hit sets it again

            // ---- 2. nearest hit -------------------------------------------------------------------------------------
            float hit_t = SP_INF;
            int hit_id = -1;
            bool outer = true;                                 // hit.orient > 0
            const uint32_t act = __ballot_sync(0xffffffffu, active);
            if (active) {
                const uint32_t mode = meta_mode(r.meta), src = meta_src(r.meta);
                if (src != SP_SRC_NONE && mode == SP_SELF_ZERO) {
                    // the ray dives back into the surface it starts on: immediate hit at t = 0 (sp_kernels.cu)
                    const DCollider& c0 = sc.colliders[src];
                    float3 Nc = to_f3(sp_collider_normal<float>(c0.type, c0.p, from_f3<float>(r.o)));
                    hit_t = 0.f; hit_id = (int)src; outer = dot(r.d, Nc) < 0.f;
                } else {
                    uint32_t bcode = 0xFFFFFFFFu;
                    sp_intersect_lean(s_dyn, r.o, r.d, self_tag, mode, hit_t, bcode);
                    if (hit_t < SP_INF) {
                        hit_id = chunk_ids[bcode & 0x7FFFFFFFu];
                        outer = (bcode & 0x80000000u) == 0u;
                    }
                    if (BIG) {
                        ChunkBest best; best.t = hit_t; best.idx = -1; best.orient = 0;
                        sp_bvh_nearest(sc.bvh, r.o, r.d, src == SP_SRC_NONE ? -1 : (int)src, mode, false, -SP_INF, best);
                        if (best.idx >= 0) { hit_t = best.t; hit_id = best.idx; outer = best.orient > 0; }
                    }
                }
                // rays traced, by depth: one shared-memory update per distinct depth in the warp
                const uint32_t depth = meta_depth(r.meta);
                const uint32_t peers = __match_any_sync(act, depth);
                if (lane == (uint32_t)(__ffs(peers) - 1)) hist[depth] += __popc(peers);
                if (L0 && a.source == SP_SRC_USER && depth == 0u) {
                    if (a.out_hit) a.out_hit[r.pix] = hit_id;
                    if (a.out_t) a.out_t[r.pix] = hit_t;
                }
            }

            // ---- 3. what the hit does ---------------------------------------------------------------------------------
            int fan_class = -1;
            uint32_t bin = 0u, ctype = 0u;                     // bin: 1 + stash bin of the hit, 0 = none
            if (hit_id >= 0) {
                const uint2 hc = BIG ? __ldg(sc.col_cls + hit_id) : sh.cls[hit_id];
                ctype = (hc.y >> 16) & 255u;
                if (HAS_DIFF) fan_class = (int)(int8_t)(hc.x >> ((r.meta >> 3) & 24u));     // byte diffuse_reflections of the fan word
                if (meta_depth(r.meta) < (hc.y & 255u)) bin = (hc.y >> 9) & 7u;
                if (hc.y & 0x100u) {                              // emissive.py:21-23
                    const float3 add = r.thr * xyz(BIG ? __ldg(sc.col_lite + hit_id) : sh.lite[hit_id]);
                    float* px = reinterpret_cast<float*>(a.accum + r.pix);
                    if (add.x != 0.f) atomicAdd(px, add.x);
                    if (add.y != 0.f) atomicAdd(px + 1, add.y);
                    if (add.z != 0.f) atomicAdd(px + 2, add.z);
                }
            }

            // ---- 4. untextured Diffuse hits (diffuse.py:25-124) ------------------------------------------------------
            // hit point, shading normal, throughput x albedo form a fan record.  A single-ray fan stays in the lane's
            // registers (the loop samples and traces it next); a many-ray fan becomes the lane's parent if its slot is
            // free (level 0), else it is queued for the next launch.
            if (HAS_DIFF) {
                bool emit = false;
                if (fan_class >= 0) {
                    const float4 lite = BIG ? __ldg(sc.col_lite + hit_id) : sh.lite[hit_id];
                    const float inv_m = (meta_dr(r.meta) < 1u) ? lite.w : 1.f;
                    const float3 thr = r.thr * xyz(lite) * inv_m;
                    if (any_nonzero(thr)) {
                        const DCollider& col = sc.colliders[hit_id];
                        const float3 P = fma3(r.d, hit_t, r.o);
                        const float3 Nc = to_f3(sp_collider_normal<float>((int)ctype, col.p, from_f3<float>(P)));
                        const float3 N = outer ? Nc : -Nc;
                        // sampled directions lie in the hemisphere of N: they leave a planar / outer surface
                        // and cross the interior of a convex collider hit from inside (sp_shade.cuh)
                        const bool planar = (ctype == SP_COLLIDER_PLANE || ctype == SP_COLLIDER_TRIANGLE);
                        const uint32_t mode = (planar || outer) ? SP_SELF_SKIP : SP_SELF_FAR;
                        r.meta = sp_pack_meta(meta_depth(r.meta) + 1u, meta_dr(r.meta) + 1u, meta_medium(r.meta), (uint32_t)hit_id, mode);
                        r.o = fma3(N, 1e-6f, P); r.d = N; r.thr = thr;
                        if (fan_class == 0) {
                            r.path = sp_child_path(r.path, 0u);
                            flags = 1u;
                        } else if (L0 && child >= slot_mult) {
                            my_rec[0] = make_float4(r.o.x, r.o.y, r.o.z, __uint_as_float(r.pix));
                            my_rec[32] = make_float4(r.d.x, r.d.y, r.d.z, __uint_as_float(r.path));
                            my_rec[64] = make_float4(r.thr.x, r.thr.y, r.thr.z, __uint_as_float(r.meta));
                            child = 0u; slot_mult = (uint32_t)sc.fan_mult[fan_class];
                        } else {
                            emit = true;
                        }
                    }
                }
                uint32_t todo = __ballot_sync(0xffffffffu, emit);
                while (todo) {                                    // one round per fan class present in the warp
                    const int c = __shfl_sync(0xffffffffu, fan_class, __ffs(todo) - 1);
                    const uint32_t bc = __ballot_sync(0xffffffffu, emit && fan_class == c);
                    todo &= ~bc;
                    const PathGrant g = spp_slab_alloc(slabs + 2 * (1 + c), __popc(bc), 1u + (uint32_t)c, sg[SPP_SEG_SLAB], a.out, lane);
                    if (emit && fan_class == c) {
                        const uint32_t slot = spp_slab_pos(g, __popc(bc & spp_lanemask_lt()));
                        if (slot != SP_SLOT_NONE) {
                            SP_ASSERT(a.out.stats, slot < a.out.fans.capacity, SP_CHK_SLOT);
                            sp_write_record(a.out.fans, slot, r.o, r.d, r.thr, r.pix, r.path, r.meta);
                        }
                    }
                }
            }

            // ---- 5. everything else: stash by material kind, shade 32 at a time ----------------------------------------
            {
                uint32_t todo = __ballot_sync(0xffffffffu, bin != 0u);
                while (todo) {
                    const uint32_t b = __shfl_sync(0xffffffffu, bin, __ffs(todo) - 1);
                    const uint32_t mb = __ballot_sync(0xffffffffu, bin == b);
                    todo &= ~mb;
                    uint32_t* const stash = stash0 + sc.stash_slot[b - 1u] * SPP_STASH_STRIDE;
                    uint32_t n_st = nst[b - 1u];
                    if (bin == b) {
                        const uint32_t j = n_st + __popc(mb & spp_lanemask_lt());
                        SP_ASSERT(a.out.stats, j < SPP_STASH_CAP, SP_CHK_STASH);
                        uint32_t* st = stash + j;
                        st[0 * SPP_STASH_CAP] = __float_as_uint(r.o.x); st[1 * SPP_STASH_CAP] = __float_as_uint(r.o.y); st[2 * SPP_STASH_CAP] = __float_as_uint(r.o.z);
                        st[3 * SPP_STASH_CAP] = __float_as_uint(r.d.x); st[4 * SPP_STASH_CAP] = __float_as_uint(r.d.y); st[5 * SPP_STASH_CAP] = __float_as_uint(r.d.z);
                        st[6 * SPP_STASH_CAP] = __float_as_uint(r.thr.x); st[7 * SPP_STASH_CAP] = __float_as_uint(r.thr.y); st[8 * SPP_STASH_CAP] = __float_as_uint(r.thr.z);
                        st[9 * SPP_STASH_CAP] = r.pix; st[10 * SPP_STASH_CAP] = r.path; st[11 * SPP_STASH_CAP] = r.meta;
                        st[12 * SPP_STASH_CAP] = __float_as_uint(hit_t);
                        st[13 * SPP_STASH_CAP] = (uint32_t)hit_id | (outer ? 0x80000000u : 0u);
                    }
                    n_st += __popc(mb);
                    __syncwarp();
                    if (n_st >= 32u) {
                        n_st -= 32u;
                        spp_flush_bin<FEAT>(b - 1u, &sc, &a, stash, slabs, s_lin_lut, n_st, 32u, sg[SPP_SEG_SLAB]);
                    }
                    __syncwarp();
                    if (lane == 0) nst[b - 1u] = n_st;
                    __syncwarp();
                }
            }
        }
    }

    // ---- drain: what is left in the stashes, then the unused tails of the slabs ----------------------------------------
    __syncwarp();
    for (uint32_t b = 0; b < SP_N_BINS; ++b) {
        const uint32_t n_st = nst[b];
        if (n_st) spp_flush_bin<FEAT>(b, &sc, &a, stash0 + sc.stash_slot[b] * SPP_STASH_STRIDE, slabs, s_lin_lut, 0u, n_st, sh.seg[0][SPP_SEG_SLAB]);
        __syncwarp();
    }
    for (uint32_t q = 0; q < SPP_N_QUEUES; ++q) {
        const uint32_t next = slabs[2 * q], end = slabs[2 * q + 1];
        const RayQueue& rq = (q == 0) ? a.out.rays : a.out.fans;
        for (uint32_t s = next + lane; s < end; s += 32u) sp_write_dead(rq, s);
    }
    for (uint32_t d = lane; d < SP_MAX_LEVELS; d += 32u) {
        const uint32_t n = hist[d];
        if (n) atomicAdd(&a.out.stats->rays[d], (unsigned long long)n);
    }
}
