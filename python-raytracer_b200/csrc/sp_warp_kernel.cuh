// Warp-autonomous level kernel for small untextured Monte-Carlo scenes (the Cornell box family).
//
// Same contract as sp_level_kernel (sp_kernels.cu): one launch consumes every ray of one recursion depth
// of get_raycolor (ray.py:122-148) from the previous level's queues and appends the next level's records.
// What differs is how the work is organised inside an SM.  sp_level_kernel parks all 512 rays of a CTA
// iteration in shared memory, regroups them by the material they hit and shades them behind a CTA barrier;
// ncu put a third of its warp instructions into that bookkeeping (shared-memory atomics, ballots, list
// handling, chunk hand-out) and 8 % of its issue cycles into the two barriers.  Here every warp is on its
// own — no barrier and no shared-memory atomic in the loop, one global atomic per 8 iterations:
//   * the cheap, common outcomes of a hit are handled on the spot: a Diffuse hit writes its fan record
//     (diffuse.py:25-124: hit point, shading normal, throughput x albedo), an Emissive hit adds
//     throughput x colour to its pixel (emissive.py:21-23), a hit that is black by construction does nothing;
//   * the expensive outcome — Refractive (refractive.py:24-123: complex Fresnel, two children) — goes into a
//     *warp-private stash* in shared memory and is shaded 32 hits at a time, so that code runs with full
//     lanes no matter how the glass hits are scattered over the rays;
//   * queue slots come from *warp-private slabs*: a warp reserves a run of slots with one global atomic
//     (per 64-256 records instead of per CTA iteration) and hands them out with ballot + popc.  A slab that
//     cannot take a whole request is finished by the first ranks and the rest go to a fresh one, so the
//     only unused slots are each warp's last slab; they are filled with dead records at exit;
//   * work items are walked segment by segment (ray records, then each fan class), so the class of an item and
//     its constants (multiplicity, magic divisor, queue base) are uniform and read from shared memory; inside a
//     segment the warps draw batches of 8 iterations from a counter (rays differ in cost; a static split left a
//     fifth of the warp slots idle at the end of every launch);
//   * the record a lane needs in its next iteration is copied into the lane's shared-memory slot with cp.async
//     while the current iteration runs;
//   * the warp's index goes through __reduce_max_sync, which puts it and everything derived from it (item cursor,
//     stash / slab addresses, counters) on the uniform datapath: 64 registers, 4 CTAs per SM, no spills to speak of.
// Eligibility (host, sp_use_warp_kernel): queue-fed level, material set Diffuse + Refractive + Emissive
// without textures, one geometry chunk, no BVH, fewer than 64 colliders, fan multiplicities <= 1024.
#pragma once
#include "sp_launch.h"
#include "sp_sampling.cuh"
#include "sp_shade.cuh"

#ifndef SPW_BLOCK
#define SPW_BLOCK 256
#endif
#define SPW_WARPS (SPW_BLOCK / 32)
#ifndef SPW_CTAS
#define SPW_CTAS 4                   // resident CTAs per SM the register allocation aims for (64 registers, no spills; measured: 3 CTAs at 80 registers -4 %, 5 at 48 -6 %, 6 at 40 -5 %)
#endif
#ifndef SPW_BATCH
#define SPW_BATCH 8                  // iterations (of 32 items) a warp draws from the work counter at a time; measured on the
#endif                               // headline frame: 1 -> 17.5, 2 -> 24.9, 4 -> 34.1, 8 -> 35.5, 16 -> 35.4 Grays/s (same-address atomics)
#define SPW_STASH_WORDS 14           // o d thr pix path meta t (id | orient)
#define SPW_STASH_CAP 64             // < 32 left over + 32 pushed
#define SPW_MAX_COLLIDERS 64           // = SP_BVH_MIN_COLLIDERS: larger scenes go through the BVH variant
#define SPW_N_QUEUES (1 + SP_MAX_FAN_CLASSES)

struct WarpShared {
    uint32_t stash[SPW_WARPS][SPW_STASH_WORDS][SPW_STASH_CAP];
    float4 rec[SPW_WARPS][3][32];                  // the record each lane reads in its next iteration (cp.async)
    uint32_t slab[SPW_WARPS][SPW_N_QUEUES][2];     // per warp and output queue: next free slot, end of the slab
    uint32_t seg[SPW_N_QUEUES][8];                 // per work-item segment: SPW_SEG_* constants
    uint2 cls[SPW_MAX_COLLIDERS];                  // per collider: what a hit does (sp_hit_class below)
    float2 src_info[SPW_MAX_COLLIDERS];            // per collider: position in the chunk's id array (as int bits), cosine-pdf weight
    float4 imp[SP_MAX_IMPORTANCE];                 // importance list (centre, radius): indexed per lane when a cap is picked
    int ids[SPW_MAX_COLLIDERS];                    // position in the chunk's id array -> collider id
    float4 lite[SPW_MAX_COLLIDERS];                // per collider: albedo / emitted colour, 1 / diffuse_rays
};

// Slots for `tot` records of output queue q, requested by the whole warp at once.  Rank x of the request
// lives at  x < rem ? first + x : fresh + (x - rem).
struct SlabGrant { uint32_t first, rem, fresh; };
SP_DEV uint32_t sp_slab_pos(const SlabGrant& g, uint32_t x) {
    if (x < g.rem) return g.first + x;
    return g.fresh == SP_SLOT_NONE ? SP_SLOT_NONE : g.fresh + (x - g.rem);
}

SP_DEV SlabGrant sp_slab_alloc(uint32_t* slab, uint32_t tot, uint32_t q, uint32_t slab_size, const LevelOut& out, uint32_t lane) {
    SlabGrant g;
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

// What a hit on a collider does, as two words the level loop can test with a few instructions:
//   x: byte d = fan class a Diffuse hit emits for a ray with diffuse_reflections == d (diffuse.py:34, 85), 0xFF = none
//   y: [0:8) max_ray_depth if the material is Refractive (refractive.py:38) else 0, bit 8 Emissive, [16:24) collider type
SP_DEV uint2 sp_hit_class(const DColInfo& ci) {
    uint32_t fan = 0xFFFFFFFFu, misc = (uint32_t)ci.type << 16;
    if (ci.kind == SP_MAT_DIFFUSE) {
        fan = 0u;
        for (int dr = 0; dr < 4; ++dr) {
            uint32_t c = 0xFFu;
            if (dr < 1) c = ci.fan_class;
            else if (dr < (int)ci.max_dr) c = 0u;
            fan |= c << (8 * dr);
        }
    } else if (ci.kind == SP_MAT_EMISSIVE) {
        misc |= 0x100u;
    } else if (ci.kind == SP_MAT_REFRACTIVE) {
        misc |= (uint32_t)min(max((int)ci.max_ray_depth, 0), 255);
    }
    return make_uint2(fan, misc);
}

// Shade the top n (<= 32) entries of the warp's stash: all Refractive hits that can still spawn children.
template <uint32_t FEAT>
__device__ __noinline__ void sp_shade_stash(const DScene* scp, const LevelArgs* ap, uint32_t* stash, uint32_t* slabs,
                                            uint32_t first, uint32_t n, uint32_t slab_size, uint32_t lane) {
    const DScene& sc = *scp;
    const LevelArgs& a = *ap;
    const bool mine = lane < n;
    Ray s;
    HitRec h;
    s.o = s.d = s.thr = v3(0.f); s.pix = s.path = s.meta = 0u; h.t = 0.f; h.id = 0; h.orient = 1;
    int n_ray = 0;
    if (mine) {
        const uint32_t j = first + lane;
        SP_ASSERT(a.out.stats, j < SPW_STASH_CAP, SP_CHK_STASH);
        const uint32_t* st = stash + j;
        s.o = v3(__uint_as_float(st[0 * SPW_STASH_CAP]), __uint_as_float(st[1 * SPW_STASH_CAP]), __uint_as_float(st[2 * SPW_STASH_CAP]));
        s.d = v3(__uint_as_float(st[3 * SPW_STASH_CAP]), __uint_as_float(st[4 * SPW_STASH_CAP]), __uint_as_float(st[5 * SPW_STASH_CAP]));
        s.thr = v3(__uint_as_float(st[6 * SPW_STASH_CAP]), __uint_as_float(st[7 * SPW_STASH_CAP]), __uint_as_float(st[8 * SPW_STASH_CAP]));
        s.pix = st[9 * SPW_STASH_CAP]; s.path = st[10 * SPW_STASH_CAP]; s.meta = st[11 * SPW_STASH_CAP];
        h.t = __uint_as_float(st[12 * SPW_STASH_CAP]);
        const uint32_t packed = st[13 * SPW_STASH_CAP];
        h.id = (int)(packed & 0x7FFFFFFFu); h.orient = (packed & 0x80000000u) ? 1 : -1;
        const float4 raw = __ldg(reinterpret_cast<const float4*>(sc.col_info + h.id));
        const DColInfo ci = *reinterpret_cast<const DColInfo*>(&raw);
        int fan_class;
        sp_child_needs(ci, meta_depth(s.meta), meta_dr(s.meta), n_ray, fan_class);
    }
    __syncwarp();                                             // the entries are free again
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t b0 = __ballot_sync(0xffffffffu, n_ray & 1), b1 = __ballot_sync(0xffffffffu, n_ray & 2);
    const uint32_t tot = __popc(b0) + 2u * __popc(b1);
    ShadeCtx ctx;
    ctx.sc = scp; ctx.out = &a.out; ctx.shadow_slot = a.shadow_slot; ctx.lin_lut = nullptr; ctx.shadow_rays = 0;
    ctx.shq = nullptr; ctx.shq_cap = 0u; ctx.shq_count = nullptr;
    ctx.ray_slot = ctx.ray_slot1 = SP_SLOT_NONE; ctx.ray_used = 0u; ctx.fan_slot = SP_SLOT_NONE;
    if (tot) {
        const SlabGrant g = sp_slab_alloc(slabs, tot, 0u, slab_size < 64u ? 64u : slab_size, a.out, lane);
        const uint32_t rank = __popc(b0 & lt_mask) + 2u * __popc(b1 & lt_mask);
        if (n_ray >= 1) ctx.ray_slot = sp_slab_pos(g, rank);
        if (n_ray >= 2) ctx.ray_slot1 = sp_slab_pos(g, rank + 1u);
    }
    if (mine) {
        const float3 add = sp_shade<FEAT>(ctx, s, h);
        sp_accum_add(a.accum + s.pix, add);
        // reserved but unused slots become dead records
        if (ctx.ray_used < 1u && n_ray >= 1 && ctx.ray_slot != SP_SLOT_NONE) sp_write_dead(a.out.rays, ctx.ray_slot);
        if (ctx.ray_used < 2u && n_ray >= 2 && ctx.ray_slot1 != SP_SLOT_NONE) sp_write_dead(a.out.rays, ctx.ray_slot1);
    }
}

// 16-byte asynchronous copy global -> shared (LDGSTS: no register staging, completes in the background)
SP_DEV void sp_cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
SP_DEV void sp_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
SP_DEV void sp_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
SP_DEV uint32_t sp_lane_id() { uint32_t r; asm("mov.u32 %0, %%laneid;" : "=r"(r)); return r; }
SP_DEV uint32_t sp_lanemask_lt() { uint32_t r; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(r)); return r; }

// per-segment constants (shared memory, read where they are needed instead of living in registers)
enum { SPW_SEG_ITEMS = 0, SPW_SEG_MULT, SPW_SEG_MAGIC_LO, SPW_SEG_MAGIC_HI, SPW_SEG_BASE, SPW_SEG_SLAB, SPW_SEG_BATCH, SPW_SEG_WORDS = 8 };

template <uint32_t FEAT>
__global__ void __launch_bounds__(SPW_BLOCK, SPW_CTAS)
sp_warp_kernel(const __grid_constant__ DScene sc, const __grid_constant__ LevelArgs a) {
    static_assert((FEAT & ~(SP_F_DIFFUSE | SP_F_REFR)) == 0u, "inline shading covers untextured Diffuse / Emissive only");
    extern __shared__ float4 s_geom[];
    __shared__ WarpShared sh;

    const uint32_t tid = threadIdx.x;

    // a queue of an earlier level overflowed: its records are incomplete (reserved but never written), the host
    // discards the chunk and renders it again in smaller pieces
    if (*reinterpret_cast<volatile const unsigned int*>(&a.out.stats->overflow) & 0xFFFFu) return;

    // ---- work items of this launch -------------------------------------------------------------------
    const uint32_t n_rays = min(__ldg(a.in_counts), a.in_rays.capacity);
    uint32_t total = n_rays;
    for (int c = 0; c < sc.n_fan_classes; ++c) total += min(__ldg(a.in_counts + 1 + c), a.in_fan_cap[c]) * (uint32_t)sc.fan_mult[c];
    if (total == 0u) return;

    sp_stage_chunk(s_geom, sc, sc.all, 0);
    for (uint32_t i = tid; i < (uint32_t)sc.n_colliders; i += SPW_BLOCK) {
        const float4 raw = __ldg(reinterpret_cast<const float4*>(sc.col_info + i));
        sh.cls[i] = sp_hit_class(*reinterpret_cast<const DColInfo*>(&raw));
        sh.src_info[i].y = reinterpret_cast<const DColInfo*>(&raw)->w_cos;
        sh.lite[i] = __ldg(sc.col_lite + i);
    }
    if (tid < (uint32_t)sc.n_importance)
        sh.imp[tid] = make_float4(sc.importance[tid].center.x, sc.importance[tid].center.y, sc.importance[tid].center.z, sc.importance[tid].radius);
    if (tid < SPW_WARPS * SPW_N_QUEUES * 2) reinterpret_cast<uint32_t*>(sh.slab)[tid] = 0u;
    if (tid <= (uint32_t)sc.n_fan_classes) {
        // slots per slab: about an eighth of what a warp can emit in this launch, so that the unused tails stay a
        // few per cent of the queue even for small launches; a power of two in [32, 256]
        uint32_t slab_size = 32u;
        const uint32_t per_warp = total / (gridDim.x * SPW_WARPS * 8u);
        while (slab_size < 256u && slab_size * 2u <= per_warp) slab_size *= 2u;
        // segment 0: explicit ray records; segment 1 + c: the children of fan class c
        uint32_t mult = 1u, n_items = n_rays, fan_base = 0u;
        if (tid > 0) {
            mult = (uint32_t)sc.fan_mult[tid - 1];
            n_items = min(__ldg(a.in_counts + tid), a.in_fan_cap[tid - 1]) * mult;
            fan_base = a.in_fan_base[tid - 1];
        }
        uint32_t* sg = sh.seg[tid];
        sg[SPW_SEG_ITEMS] = n_items; sg[SPW_SEG_MULT] = mult; sg[SPW_SEG_BASE] = fan_base;
        // item / mult == __umul64hi(item, ceil(2^64 / mult)) for 32-bit items (mult == 1 is special-cased)
        const unsigned long long magic = tid > 0 ? sc.fan_magic[tid - 1] : 0ull;
        sg[SPW_SEG_MAGIC_LO] = (uint32_t)magic; sg[SPW_SEG_MAGIC_HI] = (uint32_t)(magic >> 32);
        sg[SPW_SEG_SLAB] = slab_size;
        // iterations per draw.  Shrinking it for short segments (so that every warp gets a share) measured slower on the
        // headline frame (33.2 vs 35.3 Grays/s): warps that find a short segment drained simply move on to the next one.
        sg[SPW_SEG_BATCH] = (uint32_t)SPW_BATCH;
    }
    __syncthreads();
    {
        const GeomChunkHeader* gh = reinterpret_cast<const GeomChunkHeader*>(s_geom);
        const int n_items = gh->n_sphere + gh->n_plane + gh->n_cuboid + gh->n_tri + gh->n_aax + gh->n_aay + gh->n_aaz;
        const int* ids = reinterpret_cast<const int*>(s_geom + gh->off_ids);
        for (int k = (int)tid; k < n_items; k += SPW_BLOCK) { sh.src_info[ids[k]].x = __int_as_float(k); sh.ids[k] = ids[k]; }
    }
    __syncthreads();

    // a warp-wide reduction hands the warp's index to the uniform datapath: everything derived from it (item
    // cursor, stash / slab addresses) can then live in uniform registers instead of one copy per lane
    const uint32_t warp = __reduce_max_sync(0xffffffffu, tid >> 5);
    uint32_t* const stash = &sh.stash[warp][0][0];
    uint32_t* const slabs = &sh.slab[warp][0][0];
    uint32_t n_st = 0;                                         // entries in the stash (warp-uniform)
    uint32_t traced = 0;                                       // rays traced by this warp (warp-uniform)

    for (int seg = 0; seg <= sc.n_fan_classes; ++seg) {
        const volatile uint32_t* sg = sh.seg[seg];
        const uint32_t n_items = sg[SPW_SEG_ITEMS];
        // Record pipeline: the three vectors of the record a lane needs in iteration i + 1 are copied into the lane's
        // own shared-memory slot (cp.async, no registers) while iteration i runs; the global-memory round trip, a
        // quarter of the stall samples of the single-ray levels, is off the critical path.
        float4* const my_rec = &sh.rec[warp][0][sp_lane_id()];
        auto fetch = [&](uint32_t first) {                     // first item of the iteration the copy is for
            const uint32_t item = first + sp_lane_id();
            if (item < n_items) {
                uint32_t rec = item;
                const RayQueue* q = &a.in_rays;
                if (seg != 0) {
                    if (sg[SPW_SEG_MULT] != 1u) {
                        const unsigned long long magic = ((unsigned long long)sg[SPW_SEG_MAGIC_HI] << 32) | sg[SPW_SEG_MAGIC_LO];
                        rec = (uint32_t)__umul64hi((unsigned long long)item, magic);
                    }
                    rec += sg[SPW_SEG_BASE];
                    q = &a.in_fans;
                }
                sp_cp_async16(my_rec, q->q0 + rec); sp_cp_async16(my_rec + 32, q->q1 + rec); sp_cp_async16(my_rec + 64, q->q2 + rec);
            }
            sp_cp_async_commit();
        };
        // Work distribution: warps draw batches of up to SPW_BATCH consecutive iterations from a per-segment counter (the
        // level's queue-count block holds it, zeroed by the host with the counts).  A static split leaves a fifth of
        // the warp slots idle at the end of a launch (rays differ in cost: glass, stash shading); the counter's round
        // trip is hidden by drawing the next batch while the current one runs.
        uint32_t* const work = const_cast<uint32_t*>(a.in_counts) + SPW_N_QUEUES + seg;
        uint32_t pending = 0;                                  // lane 0: start of the batch after the current one
        const uint32_t batch = sg[SPW_SEG_BATCH];
        auto draw = [&]() { if (sp_lane_id() == 0) pending = atomicAdd(work, 32u * batch); };
        auto drawn = [&]() { return __reduce_max_sync(0xffffffffu, sp_lane_id() == 0 ? pending : 0u); };
        draw();
        uint32_t wb = drawn(), left = batch;
        if (wb < n_items) { draw(); fetch(wb); }
#pragma unroll 1
        while (wb < n_items) {
            const uint32_t lane = sp_lane_id();
            bool active = wb + lane < n_items;
            Ray r;
            r.o = r.d = r.thr = v3(0.f); r.pix = 0; r.path = 0; r.meta = 0;
            int self_tag = -1;                                 // the source collider's position in the chunk's id array
            // ---- 1. the ray of this item -----------------------------------------------------------------
            sp_cp_async_wait_all();
            float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0, q2 = make_float4(0.f, 0.f, 0.f, __uint_as_float(SP_META_DEAD));
            if (active) { q0 = my_rec[0]; q1 = my_rec[32]; q2 = my_rec[64]; }
#ifdef SP_CHECKED
            // the slot must have been refilled since it was last read: it is poisoned after every read
            if (active) {
                SP_ASSERT(a.out.stats, __float_as_uint(q1.w) != 0xDEADBEEFu || __float_as_uint(q0.w) != 0xDEADBEEFu, SP_CHK_FETCH);
                my_rec[0].w = __uint_as_float(0xDEADBEEFu); my_rec[32].w = __uint_as_float(0xDEADBEEFu);
            }
#endif
            r.meta = __float_as_uint(q2.w);
            active = active && r.meta != SP_META_DEAD;
            // first item of the next iteration: the next 32 of this batch, or the batch drawn earlier
            uint32_t next = wb + 32u;
            if (--left == 0u) {
                next = drawn(); left = sg[SPW_SEG_BATCH];
                if (next < n_items) draw();
            }
            if (next < n_items) fetch(next);                   // the slot has been read: refill it for the next iteration
            if (seg == 0) {
                if (active) {
                    r.o = xyz(q0); r.d = xyz(q1); r.thr = xyz(q2);
                    r.pix = __float_as_uint(q0.w); r.path = __float_as_uint(q1.w);
                    const uint32_t src = meta_src(r.meta);
                    if (src != SP_SRC_NONE) self_tag = __float_as_int(sh.src_info[src].x);
                }
            } else if (active) {
                // item = rec * mult + child
                const uint32_t mult = sg[SPW_SEG_MULT], item = wb + lane;
                uint32_t child = 0u;
                if (mult != 1u) {
                    const unsigned long long magic = ((unsigned long long)sg[SPW_SEG_MAGIC_HI] << 32) | sg[SPW_SEG_MAGIC_LO];
                    child = item - (uint32_t)__umul64hi((unsigned long long)item, magic) * mult;
                }
                r.o = xyz(q0); r.thr = xyz(q2);
                r.pix = __float_as_uint(q0.w);
                r.path = sp_child_path(__float_as_uint(q1.w), child);
                const float2 si = sh.src_info[meta_src(r.meta)];      // fan records always name their source
                self_tag = __float_as_int(si.x);
                const float weight = sp_sample_diffuse_with(sc, [&](int i) { return sh.imp[i]; }, r.o, xyz(q1), si.y, r.pix, r.path, r.d);
                r.thr = r.thr * weight;
                active = weight > 0.f;                         // zero-weight samples cannot contribute: not traced
            }

            // ---- 2. nearest hit over the chunk --------------------------------------------------------------
            float hit_t = SP_INF;
            int hit_id = -1;
            bool outer = true;                                 // hit.orient > 0
            uint32_t face_axis = 3u;                           // cuboid hits of the chunk walk: the slab crossed (3: not known)
            if (active) {
                const uint32_t mode = meta_mode(r.meta);
                if (self_tag >= 0 && mode == SP_SELF_ZERO) {
                    // the ray dives back into the surface it starts on: immediate hit at t = 0 (sp_kernels.cu)
                    const uint32_t src = meta_src(r.meta);
                    const DCollider& c0 = sc.colliders[src];
                    float3 Nc = to_f3(sp_collider_normal<float>(c0.type, c0.p, from_f3<float>(r.o)));
                    hit_t = 0.f; hit_id = (int)src; outer = dot(r.d, Nc) < 0.f;
                } else {
                    uint32_t bcode = 0xFFFFFFFFu;
                    sp_intersect_lean(s_geom, r.o, r.d, self_tag, mode, hit_t, bcode);
                    if (hit_t < SP_INF) {
                        hit_id = sh.ids[bcode & 0xFFu];
                        outer = (bcode & 0x80000000u) == 0u;
                        face_axis = (bcode >> 8) & 3u;
                    }
                }
            }

            traced += __popc(__ballot_sync(0xffffffffu, active));

            // ---- 3. what the hit does -------------------------------------------------------------------------
            int fan_class = -1;
            bool glass = false;
            uint32_t ctype = 0u;
            if (hit_id >= 0) {
                const uint2 hc = sh.cls[hit_id];
                ctype = (hc.y >> 16) & 255u;
                fan_class = (int)(int8_t)(hc.x >> ((r.meta >> 3) & 24u));          // byte diffuse_reflections of the fan word
                glass = meta_depth(r.meta) < (hc.y & 255u);
                if (hc.y & 0x100u) {                              // emissive.py:21-23
                    const float3 add = r.thr * xyz(sh.lite[hit_id]);
                    sp_accum_add(a.accum + r.pix, add);
                }
            }

            // ---- 4. Diffuse hits: fan record for the next level (diffuse.py:25-124) ----------------------------
            {
                uint32_t slot = SP_SLOT_NONE;
                const uint32_t b0 = __ballot_sync(0xffffffffu, fan_class == 0);     // single-ray fans: the common case
                if (b0) {
                    const SlabGrant g = sp_slab_alloc(slabs + 2, __popc(b0), 1u, sg[SPW_SEG_SLAB], a.out, lane);
                    if (fan_class == 0) slot = sp_slab_pos(g, __popc(b0 & sp_lanemask_lt()));
                }
                uint32_t todo = __ballot_sync(0xffffffffu, fan_class > 0);
                while (todo) {                                    // one round per other fan class present in the warp
                    const int c = __shfl_sync(0xffffffffu, fan_class, __ffs(todo) - 1);
                    const uint32_t bc = __ballot_sync(0xffffffffu, fan_class == c);
                    todo &= ~bc;
                    const SlabGrant g = sp_slab_alloc(slabs + 2 * (1 + c), __popc(bc), 1u + (uint32_t)c, sg[SPW_SEG_SLAB], a.out, lane);
                    if (fan_class == c) slot = sp_slab_pos(g, __popc(bc & sp_lanemask_lt()));
                }
                if (slot != SP_SLOT_NONE) {
                    const float4 lite = sh.lite[hit_id];
                    const float inv_m = (meta_dr(r.meta) < 1u) ? lite.w : 1.f;
                    const float3 thr = r.thr * xyz(lite) * inv_m;
                    if (any_nonzero(thr)) {
                        const DCollider& col = sc.colliders[hit_id];
                        const float3 P = fma3(r.d, hit_t, r.o);
                        float3 N;
                        if (ctype == SP_COLLIDER_CUBOID && face_axis < 3u) {
                            // column `axis` of the inverse basis = outward normal of the +axis face; towards the ray's side
                            const float* ib = col.p + SP_CB_INVB + face_axis;
                            N = v3(ib[0], ib[3], ib[6]);
                            if (dot(N, r.d) > 0.f) N = -N;
                        } else {
                            const float3 Nc = to_f3(sp_collider_normal<float>((int)ctype, col.p, from_f3<float>(P)));
                            N = outer ? Nc : -Nc;
                        }
                        // sampled directions lie in the hemisphere of N: they leave a planar / outer surface
                        // and cross the interior of a convex collider hit from inside (sp_shade.cuh)
                        const bool planar = (ctype == SP_COLLIDER_PLANE || ctype == SP_COLLIDER_TRIANGLE);
                        const uint32_t mode = (planar || outer) ? SP_SELF_SKIP : SP_SELF_FAR;
                        const uint32_t meta = sp_pack_meta(meta_depth(r.meta) + 1u, meta_dr(r.meta) + 1u, meta_medium(r.meta),
                                                           (uint32_t)hit_id, mode);
                        SP_ASSERT(a.out.stats, slot < a.out.fans.capacity, SP_CHK_SLOT);
                        sp_write_record(a.out.fans, slot, fma3(N, 1e-6f, P), N, thr, r.pix, r.path, meta);
                    } else {
                        sp_write_dead(a.out.fans, slot);
                    }
                }
            }

            // ---- 5. Refractive hits: stash, shade 32 at a time ---------------------------------------------------
            {
                const uint32_t bg = __ballot_sync(0xffffffffu, glass);
                if (bg) {
                    if (glass) {
                        SP_ASSERT(a.out.stats, n_st + __popc(bg & sp_lanemask_lt()) < SPW_STASH_CAP, SP_CHK_STASH);
                        uint32_t* st = stash + n_st + __popc(bg & sp_lanemask_lt());
                        st[0 * SPW_STASH_CAP] = __float_as_uint(r.o.x); st[1 * SPW_STASH_CAP] = __float_as_uint(r.o.y); st[2 * SPW_STASH_CAP] = __float_as_uint(r.o.z);
                        st[3 * SPW_STASH_CAP] = __float_as_uint(r.d.x); st[4 * SPW_STASH_CAP] = __float_as_uint(r.d.y); st[5 * SPW_STASH_CAP] = __float_as_uint(r.d.z);
                        st[6 * SPW_STASH_CAP] = __float_as_uint(r.thr.x); st[7 * SPW_STASH_CAP] = __float_as_uint(r.thr.y); st[8 * SPW_STASH_CAP] = __float_as_uint(r.thr.z);
                        st[9 * SPW_STASH_CAP] = r.pix; st[10 * SPW_STASH_CAP] = r.path; st[11 * SPW_STASH_CAP] = r.meta;
                        st[12 * SPW_STASH_CAP] = __float_as_uint(hit_t);
                        st[13 * SPW_STASH_CAP] = (uint32_t)hit_id | (outer ? 0x80000000u : 0u);
                    }
                    n_st += __popc(bg);
                    __syncwarp();
                    if (n_st >= 32u) {
                        n_st -= 32u;
                        sp_shade_stash<SP_F_REFR>(&sc, &a, stash, slabs, n_st, 32u, sg[SPW_SEG_SLAB], lane);
                    }
                }
            }
            wb = next;
        }
    }

    // ---- drain: what is left in the stash, then the unused tails of the slabs ---------------------------------
    const uint32_t lane = sp_lane_id();
    if (n_st) sp_shade_stash<SP_F_REFR>(&sc, &a, stash, slabs, 0u, n_st, sh.seg[0][SPW_SEG_SLAB], lane);
    __syncwarp();
    for (uint32_t q = 0; q < SPW_N_QUEUES; ++q) {
        const uint32_t next = slabs[2 * q], end = slabs[2 * q + 1];
        const RayQueue& rq = (q == 0) ? a.out.rays : a.out.fans;
        for (uint32_t s = next + lane; s < end; s += 32u) sp_write_dead(rq, s);
    }
    if (lane == 0 && traced) atomicAdd(&a.out.stats->rays[a.level], (unsigned long long)traced);
}
