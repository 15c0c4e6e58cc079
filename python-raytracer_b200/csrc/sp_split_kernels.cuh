// Material-sorted wavefront for the Whitted scenes (examples 1-4: textures, Glossy, Refractive, ThinFilm, sky boxes; no
// Diffuse fans, no BVH): one level of get_raycolor (ray.py:122-148) as a *hit kernel* plus one small *shade kernel* per
// material kind, instead of the fused sp_level_kernel.
//
// Why: sp_level_kernel<Whitted> is 124 KB of SASS.  ncu (profiles/r2b_example2_level_kernel.md) shows 1760 warp
// instructions per 32 primary rays of example2, a quarter of them the CTA-wide regrouping of rays by material (parking
// in shared memory, bin counts, chunk hand-out), warps stalled 2.2 cycles per issue on instruction fetch and 1.6 at
// the two barriers, 0.56 instructions issued per scheduler-cycle.  Here the regrouping is a 4-byte append to a
// per-kind item list in global memory, every kernel's code is a fraction of the fused kernel's, and nothing waits at
// a barrier:
//   sp_hit_kernel      ray of the work item (sp_item_ray: camera / caller ray or queue record), nearest hit over the
//                      staged chunk, hit record (t, collider | face) per item — the layout sp_trace_kernel uses for
//                      BVH scenes — and the item's index appended to the list of the material kind it hit (hits that
//                      are black by construction are not listed);
//   sp_shade_kernel<K> the listed items of kind K: the ray again from its item (cheaper than 48 more bytes of record
//                      per ray), the hit record, queue slots for the children, sp_shade<K>.
// Rays, hits, children and radiance are those of sp_level_kernel (same device functions); only the order of the
// float additions into the frame differs.
#pragma once
#include "sp_launch.h"
#include "sp_shade.cuh"

#define SPS_BLOCK 256
#define SPS_WARPS (SPS_BLOCK / 32)
#define SPS_KINDS 6                      // shading bins, SP_BIN_OF_KIND order: Refractive, Glossy, ThinFilm, Diffuse, SkyBox, Emissive
#define SPS_INVALID 0xFFFFFFFFu          // list entry that names no item (unused tail of a warp's last slab)

// List entries and queue slots are handed out from warp-private slabs, as in sp_warp_kernel: a warp reserves a run of
// 32-256 of them with one global atomic and serves its next requests from it (ncu on the first version, one atomic per
// warp and request: 42 % of the shade kernel's stall samples sat on that same-address atomic, 0.27 instructions issued
// per scheduler-cycle).  Rank x of a request of `tot` lives at  x < rem ? first + x : fresh + (x - rem).
SP_DEV SlabGrant sps_slab_alloc(uint32_t* slab, uint32_t tot, uint32_t* counter, uint32_t cap, uint32_t slab_size,
                                DeviceStats* stats, uint32_t lane) {
    SlabGrant g;
    uint2 st = *reinterpret_cast<const uint2*>(slab);         // x = next free entry, y = end of the slab
    g.first = st.x; g.rem = st.y - st.x; g.fresh = SP_SLOT_NONE;
    if (tot > g.rem) {                                        // warp-uniform: finish this slab, open another
        uint32_t b = 0;
        if (lane == 0) {
            b = atomicAdd(counter, slab_size);
            if (b + slab_size > cap || b + slab_size < b) { atomicOr(&stats->overflow, 1u); b = SP_SLOT_NONE; }
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
// entries per slab: about an eighth of a warp's share of `total` requests, a power of two in [64, 256]
SP_DEV uint32_t sps_slab_size(uint32_t total) {
    uint32_t s = 64u;
    const uint32_t per_warp = total / (gridDim.x * SPS_WARPS * 8u);
    while (s < 256u && s * 2u <= per_warp) s *= 2u;
    return s;
}

template <uint32_t FEAT>
__global__ void __launch_bounds__(SPS_BLOCK, 4)
sp_hit_kernel(const __grid_constant__ DScene sc, const __grid_constant__ LevelArgs a) {
    extern __shared__ float4 s_geom[];
    __shared__ uint32_t s_slab[SPS_WARPS][SPS_KINDS][2];
    if (*reinterpret_cast<volatile const unsigned int*>(&a.out.stats->overflow) & 0xFFFFu) return;
    uint32_t n_rays = 0, fan_n[SP_MAX_FAN_CLASSES], total;
#pragma unroll
    for (int c = 0; c < SP_MAX_FAN_CLASSES; ++c) fan_n[c] = 0;
    if ((FEAT & SP_F_QUEUES) && a.source == SP_SRC_QUEUES) {
        n_rays = min(__ldg(a.in_counts), a.in_rays.capacity);
        total = n_rays;
    } else {
        total = a.n_items0;
    }
    if (total == 0u) return;
    for (int i = threadIdx.x, n = __ldg(sc.all.chunk_off + 1) - __ldg(sc.all.chunk_off); i < n; i += SPS_BLOCK) s_geom[i] = __ldg(sc.all.data + i);
    if (threadIdx.x < SPS_WARPS * SPS_KINDS * 2) (&s_slab[0][0][0])[threadIdx.x] = 0u;
    __syncthreads();

    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    uint32_t* const slabs = &s_slab[tid >> 5][0][0];
    const uint32_t slab_size = sps_slab_size(total);
    uint32_t traced = 0;
    for (unsigned long long base = (unsigned long long)blockIdx.x * SPS_BLOCK; base < total; base += (unsigned long long)gridDim.x * SPS_BLOCK) {
        const uint32_t item = (uint32_t)base + tid;
        bool active = base + tid < total;
        Ray r;
        r.o = r.d = r.thr = v3(0.f); r.pix = 0; r.path = 0; r.meta = 0;
        if (active) active = sp_item_ray<FEAT>(sc, a, item, n_rays, fan_n, r);

        HitRec hit; hit.t = SP_INF; hit.id = -1; hit.orient = 0;
        if (active) {
            const uint32_t src = meta_src(r.meta), mode = meta_mode(r.meta);
            if (src != SP_SRC_NONE && mode == SP_SELF_ZERO) {
                // the ray dives back into the surface it starts on: immediate hit at t = 0 (sp_level_kernel)
                const DCollider& c0 = sc.colliders[src];
                const float3 Nc = to_f3(sp_collider_normal<float>(c0.type, c0.p, from_f3<float>(r.o)));
                hit.t = 0.f; hit.id = (int)src; hit.orient = dot(r.d, Nc) < 0.f ? 1 : -1;
            } else {
                SelfSlot self; self.sphere = self.plane = self.cuboid = self.tri = self.aa = -1; self.mode = mode;
                if (src != SP_SRC_NONE) {
                    const uint32_t where = __ldg(&sc.col_info[src].slot);
                    if ((where >> 24) == 0u) {
                        const int ty = (int)((where >> 20) & 15u), li = (int)(where & 0xFFFFFu);
                        if (ty == 0) self.sphere = li; else if (ty == 1) self.plane = li;
                        else if (ty == 2) self.cuboid = li; else if (ty == 3) self.tri = li; else self.aa = li;
                    }
                }
                ChunkBest best; best.t = SP_INF; best.idx = -1; best.orient = 0;
                sp_intersect_chunk(s_geom, r.o, r.d, self, best);
                if (best.idx >= 0) { hit.t = best.t; hit.orient = best.orient; hit.id = sp_chunk_id(s_geom, best.idx); }
            }
        }
        int bin = SPS_KINDS;                                 // nothing to shade
        if (active) {
            traced += 1u;
            if ((FEAT & SP_F_LEVEL0) && a.level == 0 && (a.out_hit || a.out_t)) {                    // per-ray outputs (sp_trace)
                const size_t oi = (a.source == SP_SRC_USER) ? (size_t)a.user_base + (size_t)item : (size_t)item;
                if (a.out_hit) a.out_hit[oi] = hit.id;
                if (a.out_t) a.out_t[oi] = hit.t;
            }
            if (hit.id >= 0) {
                const float4 raw = __ldg(reinterpret_cast<const float4*>(sc.col_info + hit.id));
                const DColInfo ci = *reinterpret_cast<const DColInfo*>(&raw);
                bin = (int)SP_BIN_OF_KIND(ci.kind);
                int n_ray, fan_class;
                sp_child_needs(ci, meta_depth(r.meta), meta_dr(r.meta), n_ray, fan_class);
                // a Refractive / ThinFilm hit past max_ray_depth is black (refractive.py:38): nothing to shade
                if (n_ray == 0 && (ci.kind == SP_MAT_DIFFUSE || ci.kind == SP_MAT_REFRACTIVE || ci.kind == SP_MAT_THINFILM)) bin = SPS_KINDS;
            }
            a.hits[item] = make_float2(hit.t, __uint_as_float((hit.id < 0 ? 0x7FFFFFFFu : (uint32_t)hit.id) | (hit.orient > 0 ? 0x80000000u : 0u)));
        }
        // append the item to the list of its kind: one round per kind present in the warp
        uint32_t todo = __ballot_sync(0xffffffffu, bin < SPS_KINDS);
        while (todo) {
            const int b = __shfl_sync(0xffffffffu, bin, __ffs(todo) - 1);
            const uint32_t mb = __ballot_sync(0xffffffffu, bin == b);
            todo &= ~mb;
            const SlabGrant g = sps_slab_alloc(slabs + 2 * b, __popc(mb), a.kind_count + b, a.kind_cap, slab_size, a.out.stats, lane);
            if (bin == b) {
                const uint32_t pos = sp_slab_pos(g, __popc(mb & ((1u << lane) - 1u)));
                if (pos != SP_SLOT_NONE) a.kind_list[(size_t)b * a.kind_cap + pos] = item;
            }
        }
    }
    // the unused tails of the warp's slabs name no item
    __syncwarp();
    for (int b = 0; b < SPS_KINDS; ++b)
        for (uint32_t p = slabs[2 * b] + lane; p < slabs[2 * b + 1]; p += 32u) a.kind_list[(size_t)b * a.kind_cap + p] = SPS_INVALID;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) traced += __shfl_down_sync(0xffffffffu, traced, o);
    if (lane == 0 && traced) atomicAdd(&a.out.stats->rays[a.level], (unsigned long long)traced);
}

// FEAT: SP_F_TEX | one material kind's feature bit | SP_F_LEVEL0 or SP_F_QUEUES (how the rays of this level are rebuilt)
template <uint32_t FEAT, int BIN>
__global__ void __launch_bounds__(SPS_BLOCK, 4)
sp_shade_kernel(const __grid_constant__ DScene sc, const __grid_constant__ LevelArgs a) {
    __shared__ float s_lin_lut[256];
    __shared__ uint32_t s_slab[SPS_WARPS][2];
    if (*reinterpret_cast<volatile const unsigned int*>(&a.out.stats->overflow) & 0xFFFFu) return;
    const uint32_t n = min(a.kind_count[BIN], a.kind_cap);
    if (n == 0u) return;
    uint32_t n_rays = 0, fan_n[SP_MAX_FAN_CLASSES];
#pragma unroll
    for (int c = 0; c < SP_MAX_FAN_CLASSES; ++c) fan_n[c] = 0;
    if ((FEAT & SP_F_QUEUES) && a.source == SP_SRC_QUEUES) n_rays = min(__ldg(a.in_counts), a.in_rays.capacity);
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    for (uint32_t i = tid; i < 256u; i += SPS_BLOCK) s_lin_lut[i] = c_decode[SP_DECODE_LINEAR][i];
    if (tid < SPS_WARPS * 2) (&s_slab[0][0])[tid] = 0u;
    __syncthreads();
    uint32_t* const slab = &s_slab[tid >> 5][0];
    const uint32_t slab_size = sps_slab_size(2u * n);
    ShadeCtx ctx;
    ctx.sc = &sc; ctx.out = &a.out; ctx.shadow_slot = a.shadow_slot; ctx.lin_lut = s_lin_lut; ctx.shadow_rays = 0;
    ctx.shq = nullptr; ctx.shq_cap = 0u; ctx.shq_count = nullptr;
    const uint32_t* list = a.kind_list + (size_t)BIN * a.kind_cap;
    const uint32_t lt_mask = (1u << lane) - 1u;

    for (uint32_t base = blockIdx.x * SPS_BLOCK; base < n; base += gridDim.x * SPS_BLOCK) {
        const uint32_t k = base + tid;
        bool mine = k < n;
        Ray s;
        s.o = s.d = s.thr = v3(0.f); s.pix = s.path = s.meta = 0u;
        HitRec h; h.t = 0.f; h.id = 0; h.orient = 1;
        int n_ray = 0;
        const uint32_t item = mine ? __ldg(list + k) : SPS_INVALID;
        mine = item != SPS_INVALID;
        if (mine) {
            mine = sp_item_ray<FEAT>(sc, a, item, n_rays, fan_n, s);
            const float2 hr = a.hits[item];
            const uint32_t code = __float_as_uint(hr.y);
            h.t = hr.x; h.id = (int)(code & 0x7FFFFFFFu); h.orient = (code & 0x80000000u) ? 1 : -1;
            const float4 raw = __ldg(reinterpret_cast<const float4*>(sc.col_info + h.id));
            const DColInfo ci = *reinterpret_cast<const DColInfo*>(&raw);
            int fan_class;
            sp_child_needs(ci, meta_depth(s.meta), meta_dr(s.meta), n_ray, fan_class);
            if (!mine) n_ray = 0;
        }
        // queue slots of the warp's children, from the warp's slab
        const uint32_t b0 = __ballot_sync(0xffffffffu, n_ray & 1), b1 = __ballot_sync(0xffffffffu, n_ray & 2);
        const uint32_t tot = __popc(b0) + 2u * __popc(b1);
        ctx.ray_slot = ctx.ray_slot1 = SP_SLOT_NONE; ctx.ray_used = 0u; ctx.fan_slot = SP_SLOT_NONE;
        if (tot) {
            const SlabGrant g = sps_slab_alloc(slab, tot, a.out.counts, a.out.rays.capacity, slab_size, a.out.stats, lane);
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
    // the unused tail of the warp's last slab becomes dead records
    __syncwarp();
    for (uint32_t p = slab[0] + lane; p < slab[1]; p += 32u) sp_write_dead(a.out.rays, p);
    unsigned long long shr = ctx.shadow_rays;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) shr += __shfl_down_sync(0xffffffffu, shr, o);
    if (lane == 0 && shr) atomicAdd(&a.out.stats->shadow_rays, shr);
}
