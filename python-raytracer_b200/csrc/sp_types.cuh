// Device-side scene description and wavefront queue layout.
#pragma once
#include "sp_math.cuh"
#include "../../include/sightpy_b200.h"   // SP_COLLIDER_*, SP_MAT_*, SP_LIGHT_* enums

// ---- geometry stream ----------------------------------------------------------------------------
// The colliders a ray must be tested against are packed, type-sorted, into 16-byte vectors and cut
// into chunks that fit the CTA's shared-memory staging buffer.  Every chunk is self-describing:
//   [GeomChunkHeader (4 x float4)] [spheres 1 v4 each] [planes 4 v4] [cuboids 5 v4] [triangles 6 v4]
//   [collider ids: int per item, same order]
// sphere   : (cx, cy, cz, r^2)
// plane    : (N.xyz, w) (C.xyz, h) (U.xyz, -) (V.xyz, -)
// cuboid   : (B0.xyz, lo.x) (B1.xyz, lo.y) (B2.xyz, lo.z) (C.xyz, hi.x) (hi.y, hi.z, -, -)
//            B = basis rows, lo/hi = box corners relative to B*C (so tests run on O - C)
// triangle : rows of the affine map to the unit triangle (Woop): (M0.xyz, t0) (M1.xyz, t1) (N.xyz, t2) with
//            M = [p2-p1, p3-p1, N]^-1, t = -M p1, so that M O + t = (u, v, signed distance to the plane)
// aa rect  : a bounded plane whose normal and both edge axes are coordinate axes (every wall of the
//            Cornell box, the examples' floors): (C.xyz, sign of the normal) (half extents along the
//            two in-plane axes in x<y<z order, -, -); one section per normal axis, ~10 instructions
//            per test instead of ~45
#define SP_CHUNK_VEC4 2048            // at most 32 KB of (dynamic) shared memory per staged chunk
#define SP_V4_SPHERE 1
#define SP_V4_PLANE 4
#define SP_V4_CUBOID 5
#define SP_V4_TRIANGLE 3
#define SP_V4_AARECT 2
#define SP_BVH_MIN_COLLIDERS 64      // smaller scenes are looped exhaustively
// stream type codes (order of the sections and of the id array inside a chunk)
enum { SP_ST_SPHERE = 0, SP_ST_PLANE = 1, SP_ST_CUBOID = 2, SP_ST_TRI = 3, SP_ST_AAX = 4, SP_ST_AAY = 5, SP_ST_AAZ = 6 };

struct GeomChunkHeader {
    int n_sphere, n_plane, n_cuboid, n_tri;
    int off_sphere, off_plane, off_cuboid, off_tri;   // in float4 units from the chunk start
    int off_ids, n_vec4, n_aax, n_aay;                // axis-aligned rectangles by normal axis ...
    int n_aaz, off_aa, pad4, pad5;                    // ... stored back to back from off_aa
};
static_assert(sizeof(GeomChunkHeader) == 64, "header is 4 float4");

struct GeomStream {
    const float4* data;        // all chunks back to back
    const int* chunk_off;      // n_chunks + 1 offsets (float4 units)
    int n_chunks;
    int n_items;
    int max_chunk_vec4;        // largest chunk (float4 units): the kernel's dynamic shared memory
    int pad;
};

// ---- full records used at shading time (one per collider / primitive / material) -------------
struct DCollider {
    int type, prim;
    float p[44];               // slots 0-39 as sp_collider.p (include/sightpy_b200.h) + derived reciprocals (sp_surface.cuh)
};

struct DPrimitive {
    int material, max_ray_depth, mc, uv_cross;
};

struct DMaterial {
    int kind, medium, normalmap_tex, color_tex;
    int aux_tex0, aux_tex1, diffuse_rays, max_dr;
    int index_h, index_w, fan_class, precise;
    float normalmap_repeat, color_repeat;
    float3 color;
    float3 n_re, n_im;
    float roughness, spec_coeff, diff_coeff;
    float thickness, noise_factor, ambient_weight, light_intensity;
};

// Everything the level kernel needs to know about the collider a ray hit before it shades it
// (one 16-byte load): material kind, recursion limits, fan class, the cosine-pdf weight.
struct DColInfo {
    uint8_t type, kind, mc, fan_class;
    int16_t max_ray_depth, max_dr;
    uint32_t slot;             // where the collider sits in the full stream: chunk[24:32) stream type[20:24) local index[0:20)
    float w_cos;
};
static_assert(sizeof(DColInfo) == 16, "one float4");

struct DTexture {
    const uint32_t* texels;    // r | g << 8 | b << 16
    int H, W, decode, pad;
};

struct DLight { int kind; float3 vec; float3 color; };
struct DImportance { float3 center; float radius; };
// absorb = 2*Im(n)*2*pi/lambda*1e9  (refractive.py:113-121)
// grey: the three channels share one real index and |Im n| <= 1e-4 |Re n| — ordinary glass.  Between two such media
// the Fresnel reflectance is evaluated once, in real arithmetic (it differs from the complex one by (Im/Re)^2 <= 1e-8).
struct DMedium { float3 re, im, absorb; int grey; };

// bounding-volume hierarchy over the small colliders of a large scene (layout: sp_geometry.cuh)
struct DBvh {
    const float4* nodes;
    const int4* items;
    const float4* data;
    int n_nodes, n_items;
};

struct DCamera {
    float3 look_from, right, up, fwd;
    float cam_w, cam_h, lens_radius, focal_distance;
    int W, H;
    // derived by the host (divisions are the dearest instructions of primary-ray generation)
    unsigned long long w_magic;        // ceil(2^64 / W): pixel / W == __umul64hi(pixel, w_magic) for 32-bit pixel indices
    float step_x, step_y;              // cam_w / (W - 1), cam_h / (H - 1): np.linspace steps (0 for a single column / row)
    float jitter_x, jitter_y;          // cam_w / W, cam_h / H
};

#define SP_MAX_FAN_CLASSES 4
#define SP_MAX_IMPORTANCE 16
#define SP_MAX_LIGHTS 8

struct DScene {
    GeomStream all, shadow;
    DBvh bvh;                          // n_nodes == 0: every collider is in the streams
    const DCollider* colliders;
    const DColInfo* col_info;
    const float4* col_lite;        // per collider: its material's plain colour, 1 / diffuse_rays (inline shading, sp_warp_kernel.cuh)
    const double* colliders_d;     // [n][44] double payloads for the precise hit path
    const DPrimitive* prims;
    const DMaterial* mats;
    const DTexture* textures;
    const DMedium* media;
    DLight lights[SP_MAX_LIGHTS];
    DImportance importance[SP_MAX_IMPORTANCE];
    DCamera cam;
    float3 ambient;
    int n_lights, n_importance, n_colliders, n_fan_classes;
    int n_shadow_casters;              // colliders of shadow-casting primitives (stream + BVH)
    float inv_n_importance;
    unsigned long long fan_magic[SP_MAX_FAN_CLASSES];   // ceil(2^64 / fan_mult): n / mult == __umul64hi(n, magic) for n < 2^32
    int fan_mult[SP_MAX_FAN_CLASSES];     // rays per fan record of each class (class 0: 1)
    int use_warp_kernel;               // option "warp_kernel" (host side only)
    int use_split;                     // option "split_kernels" (host side only)
    uint32_t seed_lo, seed_hi;
    uint32_t philox_keys[20];          // the ten Philox round key pairs of (seed_lo, seed_hi), filled per call
};

// ---- kernel specialisation ------------------------------------------------------------------------
// The level kernel is compiled for a few sets of scene features; a scene runs the smallest set that
// covers it.  Leaving out whole materials (and the double-precision texel path) shrinks the code the
// warps of an SM fight over in the instruction cache and the registers every thread must reserve.
enum : uint32_t {
    SP_F_TEX = 1u,        // image textures, normal maps, thin-film LUTs, sky boxes: uv + texel addressing (double path)
    SP_F_GLOSSY = 2u,     // Glossy (lights, shadow rays)
    SP_F_REFR = 4u,       // Refractive
    SP_F_THIN = 8u,       // ThinFilmInterference
    SP_F_DIFFUSE = 16u,   // Diffuse (fan records, importance sampling)
    SP_F_SKY = 32u,       // SkyBox / Panorama materials
    SP_F_LEVEL0 = 64u,    // sources of a level-0 launch (camera, caller rays) and its per-ray outputs
    SP_F_QUEUES = 128u,   // source of a level >= 1 launch
    SP_F_BVH = 256u,      // scenes with many colliders: BVH traversal after the staged chunk
    SP_F_MATERIALS = 63u,
};

// ---- wavefront records --------------------------------------------------------------------------
// One record = three float4 in three SoA arrays (coalesced 16-byte accesses):
//   q0 = (O.xyz, pixel)   q1 = (V.xyz, path)   q2 = (throughput.rgb, meta)
// V is the ray direction in the ray queue, and the shading normal to sample around in the fan
// queues (the consumer generates the direction: "generate + intersect" fused, so the 20 children
// of a diffuse hit never exist in memory).
// meta: depth[0:6) | diffuse_reflections[6:8) | medium[8:16) | source collider[16:30) | self mode[30:32)
#define SP_SRC_NONE 0x3FFFu
#define SP_SELF_SKIP 0u      // leaving the source surface: it cannot be hit again
#define SP_SELF_ZERO 1u      // heading back into the source surface: immediate re-hit at t = 0
#define SP_SELF_FAR  2u      // inside a convex source collider: take its far intersection

SP_DEV uint32_t sp_pack_meta(uint32_t depth, uint32_t dr, uint32_t medium, uint32_t src, uint32_t mode) {
    return (depth & 63u) | ((dr & 3u) << 6) | ((medium & 255u) << 8) | ((src & 0x3FFFu) << 16) | (mode << 30);
}
SP_DEV uint32_t meta_depth(uint32_t m) { return m & 63u; }
SP_DEV uint32_t meta_dr(uint32_t m) { return (m >> 6) & 3u; }
SP_DEV uint32_t meta_medium(uint32_t m) { return (m >> 8) & 255u; }
SP_DEV uint32_t meta_src(uint32_t m) { return (m >> 16) & 0x3FFFu; }
SP_DEV uint32_t meta_mode(uint32_t m) { return m >> 30; }

struct RayQueue {
    float4* q0; float4* q1; float4* q2;
    uint32_t capacity;
};

#define SP_MAX_LEVELS 64
// Per-level counters live in one device array so a whole chunk of levels needs a single memset:
//   counts[level][0] = ray-queue records, counts[level][1 + c] = fan-queue records of class c,
//   counts[level][1 + SP_MAX_FAN_CLASSES + k] = work counter of segment k of the launch that consumes the level
//   (sp_warp_kernel.cuh: items handed out so far)
#define SP_COUNTS_PER_LEVEL (2 * (1 + SP_MAX_FAN_CLASSES))

struct DeviceStats;
struct LevelOut {
    RayQueue rays;                         // explicit-direction records for the next level
    RayQueue fans;                         // fan records of all classes share one arena ...
    uint32_t fan_base[SP_MAX_FAN_CLASSES]; // ... cut into per-class segments
    uint32_t fan_cap[SP_MAX_FAN_CLASSES];
    uint32_t* counts;                      // counts[0] rays, counts[1 + c] fan class c (next level)
    DeviceStats* stats;
};

// Development build (-DSP_CHECKED, `make checked`): the hand-rolled protocols of the level kernels (warp-private
// stash and slabs, cp.async slot reuse, queue slot reservations) assert their invariants and report through bits
// 16+ of DeviceStats::overflow, which the host turns into an error (compute-sanitizer is not available on the pool).
#ifdef SP_CHECKED
#define SP_ASSERT(stats, cond, code) do { if (!(cond)) atomicOr(&(stats)->overflow, 0x10000u << (code)); } while (0)
#else
#define SP_ASSERT(stats, cond, code) do { } while (0)
#endif
enum { SP_CHK_SLOT = 0, SP_CHK_STASH = 1, SP_CHK_SLAB = 2, SP_CHK_FETCH = 3, SP_CHK_BIN = 4 };

struct DeviceStats {
    unsigned long long rays[SP_MAX_LEVELS];
    unsigned long long shadow_rays;
    // -DSP_PHASE_TIMING builds: warp-cycles spent in [0] ray generation, [1] intersection, [2] park + count,
    // [3] waiting at barrier A, [4] shading, [5] waiting at barrier C (summed over warps and launches)
    unsigned long long phase_cycles[6];
    unsigned int overflow;             // bit 0: a queue was too small; bits 16+: SP_ASSERT failures (SP_CHK_*)
    unsigned int pad;
};
