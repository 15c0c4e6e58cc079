// Material evaluation for one hit + emission of secondary rays into the next wavefront level.
//
// Every reference ``Material.get_color`` is linear in the radiance of the rays it spawns
// (colour = local term + sum_i w_i * L(child_i)), so the recursion tree of get_raycolor
// (ray.py:122-148) flattens exactly: a ray carries the product of the weights on its path
// ("throughput"), adds  throughput * local term  to its pixel, and hands  throughput * w_i  to its
// children.  Restates glossy.py:25-110, refractive.py:24-123, thin_film_interference.py:24-115,
// diffuse.py:25-124 (its sampling half lives in sp_sampling.cuh), emissive.py:21-23,
// skybox.py:51-94, material.py:18-36, lights.py:25-52.
#pragma once
#include "sp_geometry.cuh"
#include "sp_rng.cuh"
#include "sp_surface.cuh"


__constant__ float c_decode[2][256];      // texel byte -> value (SP_DECODE_PLAIN / SP_DECODE_LINEAR)

struct Ray {
    float3 o, d, thr;
    uint32_t pix, path, meta;
};

// Add a ray's radiance to its pixel of the accumulation frame: one 16-byte reduction (red.global.add.v4.f32, sm_90+)
// instead of three scalar ones — a third of the L2 atomic traffic and of the instructions around it.
#ifndef SP_VECTOR_RED
#define SP_VECTOR_RED 1
#endif
SP_DEV void sp_accum_add(float4* px, float3 add) {
#if SP_VECTOR_RED
    if (any_nonzero(add))
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(px), "f"(add.x), "f"(add.y), "f"(add.z), "f"(0.f) : "memory");
#else
    float* f = reinterpret_cast<float*>(px);
    if (add.x != 0.f) atomicAdd(f, add.x);
    if (add.y != 0.f) atomicAdd(f + 1, add.y);
    if (add.z != 0.f) atomicAdd(f + 2, add.z);
#endif
}


// ---- queue records ---------------------------------------------------------------------------------
// Slots are reserved per CTA *before* shading from an upper bound of what each hit can emit
// (sp_child_needs); a reserved slot whose child is not produced after all (zero throughput, total
// internal reflection) is filled with a dead record that the consumer skips.
#define SP_SLOT_NONE 0xFFFFFFFFu
#define SP_META_DEAD 0xFFFFFFFFu

SP_DEV void sp_write_record(const RayQueue& q, uint32_t slot, float3 o, float3 v, float3 thr, uint32_t pix,
                            uint32_t path, uint32_t meta) {
    q.q0[slot] = make_float4(o.x, o.y, o.z, __uint_as_float(pix));
    q.q1[slot] = make_float4(v.x, v.y, v.z, __uint_as_float(path));
    q.q2[slot] = make_float4(thr.x, thr.y, thr.z, __uint_as_float(meta));
}
SP_DEV void sp_write_dead(const RayQueue& q, uint32_t slot) {
    q.q2[slot] = make_float4(0.f, 0.f, 0.f, __uint_as_float(SP_META_DEAD));
}

// Upper bound of the records a hit emits: n_ray explicit rays (0..2) and at most one fan record of
// class fan_class (-1 = none).  Must stay in step with sp_shade below.
SP_DEV void sp_child_needs(const DColInfo& ci, uint32_t depth, uint32_t dr, int& n_ray, int& fan_class) {
    n_ray = 0; fan_class = -1;
    const bool alive = (int)depth < (int)ci.max_ray_depth;
    switch (ci.kind) {
    case SP_MAT_GLOSSY: n_ray = alive ? 1 : 0; break;
    case SP_MAT_REFRACTIVE: n_ray = alive ? (ci.mc ? 1 : 2) : 0; break;
    case SP_MAT_THINFILM: n_ray = alive ? 2 : 0; break;
    case SP_MAT_DIFFUSE:
        if (dr < 1u) fan_class = ci.fan_class;
        else if ((int)dr < (int)ci.max_dr) fan_class = 0;
        break;
    default: break;
    }
}

// Texel bytes -> values.  SP_DECODE_PLAIN is byte / 256 (exact in float); SP_DECODE_LINEAR goes through the
// 256-entry sRGB table, read from a shared-memory copy: the lanes of a warp index it with unrelated bytes,
// which constant memory would serialise.
SP_DEV float3 sp_fetch_texel(const DTexture& t, int offset, const float* __restrict__ lin_lut) {
    const uint32_t w = __ldg(t.texels + offset);
    const uint32_t r = w & 255u, g = (w >> 8) & 255u, b = (w >> 16) & 255u;
    if (t.decode == SP_DECODE_PLAIN) return v3((float)r, (float)g, (float)b) * (1.f / 256.f);
    return v3(lin_lut[r], lin_lut[g], lin_lut[b]);
}

// ---- hit geometry in float or double ----------------------------------------------------------------
struct HitGeom {
    float3 P;            // hit point
    float3 Nc;           // collider normal (outward, un-oriented)
    float cos_i;         // -D . (Nc * orientation)
    float t;             // distance travelled
    int off_color, off_normal, off_aux0, off_aux1;   // texel offsets (valid where the texture exists)
};

template <typename T, uint32_t FEAT>
SP_DEV HitGeom sp_eval_hit(const DScene& sc, const DMaterial& m, const DPrimitive& prim, const T* cp, int ctype,
                           const Ray& r, const HitRec& h) {
    HitGeom g;
    tv3<T> O = from_f3<T>(r.o), D = from_f3<T>(r.d);
    T t = (T)h.t;
    if (sizeof(T) == 8 && h.t > 0.f) t = sp_refine_t<T>(ctype, cp, O, D, h.orient, t);
    tv3<T> P = O + D * t;
    tv3<T> Nc = sp_collider_normal<T>(ctype, cp, P);
    g.P = to_f3(P); g.Nc = to_f3(Nc); g.t = (float)t;
    g.cos_i = (float)(-(tdot(D, Nc)) * (T)h.orient);
    g.off_color = g.off_normal = g.off_aux0 = g.off_aux1 = 0;
    bool need_uv = (m.color_tex >= 0) || (m.normalmap_tex >= 0) || (m.kind == SP_MAT_THINFILM);
    if ((FEAT & SP_F_TEX) && need_uv) {
        T u, v;
        sp_collider_uv<T>(ctype, cp, P, Nc, prim.uv_cross != 0, u, v);
        if (m.color_tex >= 0) {
            const DTexture& tx = sc.textures[m.color_tex];
            g.off_color = sp_texel_offset<T>(u, v, tx.H, tx.W, (T)m.color_repeat, tx.H, tx.W);
        }
        if (m.normalmap_tex >= 0) {
            const DTexture& tx = sc.textures[m.normalmap_tex];
            g.off_normal = sp_texel_offset<T>(u, v, tx.H, tx.W, (T)m.normalmap_repeat, tx.H, tx.W);
        }
        if ((FEAT & SP_F_SKY) && m.kind == SP_MAT_SKYBOX && m.aux_tex0 >= 0) {       // lightmap indexed with the env shape
            const DTexture& tx = sc.textures[m.aux_tex0];
            g.off_aux0 = sp_texel_offset<T>(u, v, m.index_h, m.index_w, (T)m.color_repeat, tx.H, tx.W);
        }
        if ((FEAT & SP_F_THIN) && m.kind == SP_MAT_THINFILM) {
            // thickness noise, then the reflectance LUT row/column (thin_film_interference.py:46-72)
            T thick = (T)m.thickness;
            if (m.noise_factor != 0.f) {
                const DTexture& nz = sc.textures[m.aux_tex1];
                int off = sp_texel_offset<T>(u, v, nz.H, nz.W, (T)0.5, nz.H, nz.W);
                T nval = (T)(__ldg(nz.texels + off) & 255u) * (T)(1.0 / 256.0);
                thick = thick + (T)m.noise_factor * (nval - (T)0.5);
            }
            const DTexture& lut = sc.textures[m.aux_tex0];
            T ci = -(tdot(D, Nc)) * (T)h.orient;
            long long row = (long long)(ci * (T)lut.H), colm = (long long)thick;
            if (row < 0) row += lut.H;                              // numpy negative indices wrap
            if (colm < 0) colm += lut.W;
            row = row < 0 ? 0 : (row >= lut.H ? lut.H - 1 : row);   // (reference: IndexError)
            colm = colm < 0 ? 0 : (colm >= lut.W ? lut.W - 1 : colm);
            g.off_aux0 = (int)(row * lut.W + colm);
        }
    }
    return g;
}

// how a child ray leaves the collider it starts on (see sp_geometry.cuh header)
//   side   : +1 origin displaced to the outward-normal side, -1 to the other side
//   heading: D . Nc
SP_DEV uint32_t sp_self_mode(int ctype, float side, float heading, int& zero_orient) {
    zero_orient = heading < 0.f ? 1 : -1;
    if (ctype == SP_COLLIDER_PLANE || ctype == SP_COLLIDER_TRIANGLE)
        return (side * heading >= 0.f) ? SP_SELF_SKIP : SP_SELF_ZERO;
    if (side > 0.f) return heading >= 0.f ? SP_SELF_SKIP : SP_SELF_ZERO;
    return heading < 0.f ? SP_SELF_FAR : SP_SELF_ZERO;
}

SP_DEV float3 sp_schlick(float3 f0, float c) {
    float w = 1.f - c;
    float w5 = w * w * w * w * w;
    return f0 + (v3(1.f) - f0) * w5;
}

// F0 = |(n1 - n2) / (n1 + n2)|^2 per channel, complex indices (glossy.py:66, 91)
SP_DEV float3 sp_f0(float3 re1, float3 im1, float3 re2, float3 im2) {
    auto ch = [](float r1, float i1, float r2, float i2) {
        return cabs2(cx(r1 - r2, i1 - i2)) / cabs2(cx(r1 + r2, i1 + i2));
    };
    return v3(ch(re1.x, im1.x, re2.x, im2.x), ch(re1.y, im1.y, re2.y, im2.y), ch(re1.z, im1.z, re2.z, im2.z));
}

// Nearest shadow-caster distance from `o` along `d` (glossy.py:53-57); walks the shadow stream
// straight from global memory / L1 (uniform addresses across the warp).
template <uint32_t FEAT>
SP_DEV float sp_shadow_nearest(const DScene& sc, float3 o, float3 d, int src_id, uint32_t mode, float dist,
                               const int2* __restrict__ shadow_slot) {
    ChunkBest best; best.t = SP_INF; best.idx = -1; best.orient = 0;
    int2 where = (src_id >= 0 && shadow_slot) ? __ldg(shadow_slot + src_id) : make_int2(-1, -1);
    for (int c = 0; c < sc.shadow.n_chunks; ++c) {
        const float4* ch = sc.shadow.data + __ldg(sc.shadow.chunk_off + c);
        SelfSlot self; self.sphere = self.plane = self.cuboid = self.tri = self.aa = -1; self.mode = mode;
        if (where.x == c) {
            int ty = where.y >> 28, li = where.y & 0x0FFFFFFF;
            if (ty == 0) self.sphere = li; else if (ty == 1) self.plane = li;
            else if (ty == 2) self.cuboid = li; else if (ty == 3) self.tri = li; else self.aa = li;
        }
        sp_intersect_chunk(ch, o, d, self, best);
    }
    if ((FEAT & SP_F_BVH) && best.t >= dist) sp_bvh_nearest(sc.bvh, o, d, src_id, mode, true, dist, best);
    return best.t;
}

struct ShadeCtx {
    const DScene* sc;
    const LevelOut* out;       // kernel parameter (constant bank), not copied into registers
    const int2* shadow_slot;    // collider id -> (chunk, type << 28 | local index) in the shadow-caster stream
    const float* lin_lut;       // shared-memory copy of the sRGB -> linear table
    unsigned long long shadow_rays;
    float4* shq; uint32_t shq_cap; uint32_t* shq_count;      // deferred shadow rays (LevelArgs::shq), nullptr = traverse inline
    // slots reserved for the hit being shaded
    uint32_t ray_slot, ray_slot1, ray_used, fan_slot;   // ray_slot1: where a second child ray goes
};

SP_DEV void sp_emit_ray(ShadeCtx& cx_, const Ray& r, float3 o, float3 d, float3 thr, uint32_t k,
                        uint32_t medium, uint32_t dr, int src, uint32_t mode) {
    if (!any_nonzero(thr)) return;          // zero-weight children cannot contribute
    const uint32_t slot = cx_.ray_used == 0u ? cx_.ray_slot : cx_.ray_slot1;
    if (slot == SP_SLOT_NONE) return;       // the CTA's reservation overflowed the queue (reported to the host)
    uint32_t meta = sp_pack_meta(meta_depth(r.meta) + 1u, dr, medium, (uint32_t)src, mode);
    SP_ASSERT(cx_.out->stats, slot < cx_.out->rays.capacity, SP_CHK_SLOT);
    sp_write_record(cx_.out->rays, slot, o, d, thr, r.pix, sp_child_path(r.path, k), meta);
    cx_.ray_used += 1u;
}

// Shade one hit.  Returns the radiance to add to the ray's pixel (already times throughput).
template <uint32_t FEAT>
SP_DEV float3 sp_shade(ShadeCtx& cx_, const Ray& r, const HitRec& h) {
    const DScene& sc = *cx_.sc;
    const DCollider& col = sc.colliders[h.id];
    const int ctype = col.type;
    const DPrimitive prim = sc.prims[col.prim];
    const DMaterial& m = sc.mats[prim.material];
    const uint32_t depth = meta_depth(r.meta), dr = meta_dr(r.meta), medium = meta_medium(r.meta);

    HitGeom g = ((FEAT & SP_F_TEX) && m.precise)
                    ? sp_eval_hit<double, FEAT>(sc, m, prim, sc.colliders_d + (size_t)h.id * SP_DEV_PAYLOAD, ctype, r, h)
                    : sp_eval_hit<float, FEAT>(sc, m, prim, col.p, ctype, r, h);
    const float orient = (float)h.orient;

    // ---- terminal materials ---------------------------------------------------------------------
    if (m.kind == SP_MAT_EMISSIVE) {                                     // emissive.py:21-23
        float3 c = ((FEAT & SP_F_TEX) && m.color_tex >= 0) ? sp_fetch_texel(sc.textures[m.color_tex], g.off_color, cx_.lin_lut) : m.color;
        return r.thr * c;
    }
    if ((FEAT & SP_F_SKY) && m.kind == SP_MAT_SKYBOX) {                                   // skybox.py:51-94
        float3 c = sp_fetch_texel(sc.textures[m.color_tex], g.off_color, cx_.lin_lut);
        if (depth != 0u && m.light_intensity != 0.f)
            c += sp_fetch_texel(sc.textures[m.aux_tex0], g.off_aux0, cx_.lin_lut) * m.light_intensity;
        return r.thr * c;
    }

    // ---- shading normal (material.py:18-36) ----------------------------------------------------------
    float3 N;
    if ((FEAT & SP_F_TEX) && m.normalmap_tex >= 0) {
        float3 tx = sp_fetch_texel(sc.textures[m.normalmap_tex], g.off_normal, cx_.lin_lut);
        float3 nm = (tx - v3(0.5f)) * 2.f;
        const float* ib = col.p + (ctype == SP_COLLIDER_PLANE ? SP_PL_INVB : SP_CB_INVB);
        N = normalize0(mat3_mul(ib, nm)) * orient;
    } else {
        N = g.Nc * orient;
    }
    const float side_plus = dot(N, g.Nc) >= 0.f ? 1.f : -1.f;            // side of P + N*eps
    const float3 V = -r.d;
    const float3 nudged = fma3(N, 1e-6f, g.P);
    float3 add = v3(0.f);
    int zo;

    if ((FEAT & SP_F_DIFFUSE) && m.kind == SP_MAT_DIFFUSE) {                                // diffuse.py:25-124
        float3 diff = ((FEAT & SP_F_TEX) && m.color_tex >= 0) ? sp_fetch_texel(sc.textures[m.color_tex], g.off_color, cx_.lin_lut) : m.color;
        int cls = -1; float inv_m = 1.f;
        if (dr < 1u) { cls = m.fan_class; inv_m = 1.f / (float)m.diffuse_rays; }
        else if ((int)dr < m.max_dr) { cls = 0; }
        if (cls >= 0) {
            float3 thr = r.thr * diff * inv_m;
            if (any_nonzero(thr)) {
                // sampled directions lie in the hemisphere of N: they leave a planar / outer surface
                // and cross the interior of a convex collider hit from inside
                bool planar = (ctype == SP_COLLIDER_PLANE || ctype == SP_COLLIDER_TRIANGLE);
                uint32_t mode = (planar || side_plus > 0.f) ? SP_SELF_SKIP : SP_SELF_FAR;
                uint32_t meta = sp_pack_meta(depth + 1u, dr + 1u, medium, (uint32_t)h.id, mode);
                if (cx_.fan_slot != SP_SLOT_NONE) {
                    SP_ASSERT(cx_.out->stats, cx_.fan_slot < cx_.out->fans.capacity, SP_CHK_SLOT);
                    sp_write_record(cx_.out->fans, cx_.fan_slot, nudged, N, thr, r.pix, r.path, meta);
                    cx_.fan_slot = SP_SLOT_NONE;            // consumed
                }
            }
        }
        return add;
    }

    if ((FEAT & SP_F_GLOSSY) && m.kind == SP_MAT_GLOSSY) {                                // glossy.py:25-110
        float3 diff = (((FEAT & SP_F_TEX) && m.color_tex >= 0) ? sp_fetch_texel(sc.textures[m.color_tex], g.off_color, cx_.lin_lut) : m.color)
                      * m.diff_coeff;
        float3 color = sc.ambient * diff;
        const DMedium med = sc.media[medium];
        for (int li = 0; li < sc.n_lights; ++li) {
            const DLight& lt = sc.lights[li];
            float3 L; float dist; float3 lv; float NdotL;
            if (lt.kind == SP_LIGHT_DIRECTIONAL) {
                L = lt.vec; dist = 1.0e6f;
                NdotL = fmaxf(dot(N, L), 0.f);
                lv = lt.color * NdotL;
            } else {
                float3 to_l = lt.vec - g.P;
                dist = sqrtf(dot(to_l, to_l));
                L = to_l * (1.f / dist);
                NdotL = fmaxf(dot(N, L), 0.f);
                lv = lt.color * (NdotL / (dist * dist) * 100.f);
            }
            if (NdotL <= 0.f) continue;                // lv == 0: both light terms vanish
            // what the light adds if it is visible: Lambert + Cook-Torrance (glossy.py:59-84)
            auto lit = [&]() {
                float3 c = diff * lv;
                if (m.roughness != 0.f) {
                    float3 Hv = normalize0(L + V);
                    float3 F = sp_schlick(sp_f0(med.re, med.im, m.n_re, m.n_im), clamp01(dot(V, Hv)));
                    float a = 2.f / (m.roughness * m.roughness) - 2.f;
                    // Phong lobe x^a: exp2(a log2 x) through the MUFU pipe (relative error ~ a * 2e-7; x^0 = 1 also at x = 0)
                    float Dp = (a == 0.f ? 1.f : __powf(clamp01(dot(N, Hv)), a)) * (a + 2.f) * (0.5f / SP_PI);
                    float denom = 4.f * fminf(fmaxf(dot(N, V) * NdotL, 0.001f), 1.f);
                    c += F * lv * (Dp / denom * m.spec_coeff);
                }
                return c;
            };
            if (sc.n_shadow_casters > 0) {
                uint32_t mode = sp_self_mode(ctype, side_plus, dot(L, g.Nc), zo);
                cx_.shadow_rays++;
                if (mode == SP_SELF_ZERO) continue;        // the shadow ray starts inside its own surface: occluded
                if ((FEAT & SP_F_BVH) && cx_.shq) {
                    // queue the shadow ray for sp_shadow_kernel (one atomic per warp and light)
                    const uint32_t peers = __activemask(), lane = threadIdx.x & 31u;
                    const int leader = __ffs(peers) - 1;
                    uint32_t base = 0;
                    if ((int)lane == leader) base = atomicAdd(cx_.shq_count, (uint32_t)__popc(peers));
                    base = __shfl_sync(peers, base, leader) + __popc(peers & ((1u << lane) - 1u));
                    if (base < cx_.shq_cap) {
                        const float3 c = r.thr * lit();
                        cx_.shq[3 * (size_t)base] = make_float4(nudged.x, nudged.y, nudged.z, dist);
                        cx_.shq[3 * (size_t)base + 1] = make_float4(L.x, L.y, L.z, __uint_as_float(r.pix));
                        cx_.shq[3 * (size_t)base + 2] = make_float4(c.x, c.y, c.z, __uint_as_float(((uint32_t)h.id << 2) | mode));
                        continue;
                    }
                }
                float nearest = sp_shadow_nearest<FEAT>(sc, nudged, L, h.id, mode, dist, cx_.shadow_slot);
                if (nearest < dist) continue;
            }
            color += lit();
        }
        add = r.thr * color;
        if ((int)depth < prim.max_ray_depth) {
            const DMedium air = sc.media[0];
            float3 F = sp_schlick(sp_f0(air.re, air.im, m.n_re, m.n_im), clamp01(dot(V, N)));
            float3 R = normalize0(fma3(N, -2.f * dot(r.d, N), r.d));
            uint32_t mode = sp_self_mode(ctype, side_plus, dot(R, g.Nc), zo);
            sp_emit_ray(cx_, r, nudged, R, r.thr * F, 0u, medium, dr, h.id, mode);
        }
        return add;
    }

    if (!(FEAT & (SP_F_REFR | SP_F_THIN))) return add;
    if ((int)depth >= prim.max_ray_depth) return add;                   // refractive.py:38, thin_film...py:34

    float3 R = normalize0(fma3(N, -2.f * dot(r.d, N), r.d));
    const float3 nudged_in = fma3(N, -1e-6f, g.P);
    const uint32_t mode_refl = sp_self_mode(ctype, side_plus, dot(R, g.Nc), zo);

    if ((FEAT & SP_F_THIN) && m.kind == SP_MAT_THINFILM) {                               // thin_film_interference.py:24-115
        float3 F = sp_fetch_texel(sc.textures[m.aux_tex0], g.off_aux0, cx_.lin_lut);
        add = r.thr * sc.ambient * F;
        sp_emit_ray(cx_, r, nudged, R, r.thr * F, 0u, medium, dr, h.id, mode_refl);
        uint32_t mode_t = sp_self_mode(ctype, -side_plus, dot(r.d, g.Nc), zo);
        sp_emit_ray(cx_, r, nudged_in, r.d, r.thr * (v3(1.f) - F), 1u, medium, dr, h.id, mode_t);
        return add;
    }

    // ---- Refractive (refractive.py:24-123) ----------------------------------------------------------
    if (FEAT & SP_F_REFR) {
        const DMedium n1 = sc.media[medium];
        const uint32_t med2 = h.orient > 0 ? (uint32_t)m.medium : 0u;
        const DMedium n2 = sc.media[med2];
        float cos_i = dot(V, N);
        float s2 = 1.f - cos_i * cos_i;
        float Fc[3];
        const float re1[3] = {n1.re.x, n1.re.y, n1.re.z}, im1[3] = {n1.im.x, n1.im.y, n1.im.z};
        const float re2[3] = {n2.re.x, n2.re.y, n2.re.z}, im2[3] = {n2.im.x, n2.im.y, n2.im.z};
        if (n1.grey && n2.grey) {
            // real indices, one value for the three channels: cos_t is real, or imaginary under total internal
            // reflection, where |r_per| = |r_par| = 1 (the complex evaluation below gives the same to ~1e-8)
            const float a_ = re1[0], b_ = re2[0];
            const float ratio = __fdividef(a_, b_);
            const float c2 = 1.f - ratio * ratio * s2;
            float Fr = 1.f;
            if (c2 >= 0.f) {
                const float cos_t = sqrtf(c2);
                const float aci = a_ * cos_i, bct = b_ * cos_t, act = a_ * cos_t, bci = b_ * cos_i;
                const float per = __fdividef(aci - bct, aci + bct), par = __fdividef(act - bci, act + bci);
                Fr = 0.5f * (per * per + par * par);
            }
            Fc[0] = Fc[1] = Fc[2] = Fr;
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                cplx a = cx(re1[c], im1[c]), b = cx(re2[c], im2[c]);
                cplx ratio = a / b;
                cplx cos_t = csqrt(cx(1.f) - ratio * ratio * s2);
                // |r|^2 = |numerator|^2 / |denominator|^2: no complex division needed
                cplx aci = a * cos_i, bct = b * cos_t, act = a * cos_t, bci = b * cos_i;
                float per = __fdividef(cabs2(aci - bct), cabs2(aci + bct));
                float par = __fdividef(cabs2(act - bci), cabs2(act + bci));
                Fc[c] = 0.5f * (per + par);
            }
        }
        float3 F = v3(Fc[0], Fc[1], Fc[2]);
        float eta = (__fdividef(re1[0], re2[0]) + __fdividef(re1[1], re2[1]) + __fdividef(re1[2], re2[2])) * (1.f / 3.f);
        float sin2_t = eta * eta * s2;
        bool non_tir = sin2_t <= 1.f;
        float3 Tdir = normalize0(fma3(N, eta * cos_i - sqrtf(1.f - clamp01(sin2_t)), r.d * eta));
        float3 absorb = v3(expf(-n1.absorb.x * g.t), expf(-n1.absorb.y * g.t), expf(-n1.absorb.z * g.t));
        float3 thr = r.thr * absorb;
        uint32_t mode_t = sp_self_mode(ctype, -side_plus, dot(Tdir, g.Nc), zo);
        if (prim.mc) {
            float u[4];
            sp_draw4_keys(r.pix, r.path, SP_BLOCK_MATERIAL, sc.philox_keys, u);
            bool pick = (u[0] > (F.x + F.y + F.z) / 3.f) && non_tir;
            if (pick) sp_emit_ray(cx_, r, nudged_in, Tdir, thr, 0u, med2, dr, h.id, mode_t);
            else      sp_emit_ray(cx_, r, nudged, R, thr, 0u, medium, dr, h.id, mode_refl);
        } else {
            sp_emit_ray(cx_, r, nudged, R, thr * F, 0u, medium, dr, h.id, mode_refl);
            if (non_tir) sp_emit_ray(cx_, r, nudged_in, Tdir, thr * (v3(1.f) - F), 1u, med2, dr, h.id, mode_t);
        }
    }
    return add;
}
