// Host-callable launchers of the kernels in sp_kernels.cu (the runtime in sp_api.cpp is plain C++).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "sp_types.cuh"

// What a level launch reads its work items from.
enum { SP_SRC_CAMERA = 0, SP_SRC_USER = 1, SP_SRC_QUEUES = 2 };
// What a level-0 launch does with its rays.
enum { SP_RUN_FULL = 0, SP_RUN_DISTANCES = 1, SP_RUN_DUMP_RAYS = 2 };


struct LevelArgs {
    int level, source, run;
    // level 0, camera: item i -> pixel pix_begin + i % n_pix, sample sample_begin + i / n_pix
    // level 0, user rays: item i -> ray user_base + i (interleaved xyz), pixel id = ray index
    uint32_t pix_begin, n_pix, sample_begin, n_items0, user_base;
    unsigned long long n_pix_magic;   // ceil(2^64 / n_pix): item / n_pix == __umul64hi(item, n_pix_magic)
    // tile-sharded frames (sp_render_tiles): the "pixel" index above runs over the texels of a list of square tiles,
    // tile_size^2 per tile, row-major inside a tile; texels outside the frame (edge tiles) are skipped
    const uint32_t* tiles;         // tile ids (row-major over the frame's tile grid), nullptr = plain pixel indices
    uint32_t tile_shift, tiles_x;  // tile_size = 1 << tile_shift; tiles per row of the frame
    const float* user_o;
    const float* user_d;
    // level >= 1: records written by the previous level
    RayQueue in_rays, in_fans;
    uint32_t in_fan_base[SP_MAX_FAN_CLASSES], in_fan_cap[SP_MAX_FAN_CLASSES];
    const uint32_t* in_counts;
    LevelOut out;
    float4* accum;                 // per pixel (or per user ray): xyz = sum of radiance
    // optional per-item outputs of level 0
    int32_t* out_hit; float* out_t; float* out_o; float* out_d;
    float* out_n;                  // oriented collider normal at the primary hit, 3 floats per item (sp_aovs)
    const int2* shadow_slot;
    // BVH scenes: nearest hits found ahead of the level launch by sp_trace_kernel, one per work item:
    // (t, collider id | outer face << 31; id 0x7FFFFFFF = miss).  Used when the launch has at most hits_cap items.
    float2* hits;
    uint32_t hits_cap;
    // BVH scenes: shadow rays of Glossy hits (glossy.py:53-57) are not traversed inside the shading phase but queued
    // — origin | distance to the light, direction | pixel, radiance the light adds if it is visible | source collider
    // and self mode — and answered by sp_shadow_kernel after the level launch.  shq_count[0]: requests queued,
    // shq_count[1]: that kernel's work counter.  A request that finds the queue full is traversed on the spot.
    float4* shq;
    uint32_t shq_cap;
    uint32_t* shq_count;
    // Whitted scenes, material-sorted wavefront (sp_split_kernels.cuh): per shading bin the items that hit that material
    // kind (bin b at kind_list + b * kind_cap, kind_count[b] entries); hit records go to `hits` above.
    uint32_t* kind_list;
    uint32_t kind_cap;
    uint32_t* kind_count;
};

struct ResolveArgs {
    const float4* accum;
    uint32_t n_pix;
    double spp;                    // divide the sums by this (scene.py:119)
    float* out_linear;             // 3 planes of n_pix floats (nullable)
    uint8_t* out_srgb8;            // n_pix * 3 interleaved (nullable)
};

uint32_t sp_pick_material_set(uint32_t needed_features);          // smallest compiled kernel variant covering them
bool sp_use_warp_kernel(const DScene& sc, uint32_t material_set);      // queue-fed levels run sp_warp_kernel (sp_warp_kernel.cuh)
int sp_level_grid(int device, const DScene& sc, uint32_t material_set, bool level0);   // CTAs of a persistent launch
cudaError_t sp_launch_level(const DScene& sc, const LevelArgs& a, uint32_t material_set, int grid, cudaStream_t st);
bool sp_can_pretrace(const DScene& sc, uint32_t material_set);           // scene behind a BVH with one staged chunk
cudaError_t sp_launch_shadow(const DScene& sc, const LevelArgs& a, int device, cudaStream_t st);
cudaError_t sp_launch_trace(const DScene& sc, const LevelArgs& a, uint32_t material_set, int device, cudaStream_t st);
// Whitted scenes without a BVH: a level as sp_hit_kernel + one sp_shade_kernel per material kind present (kind_mask: bit
// SP_MAT_*), see sp_split_kernels.cuh.  Returns the number of kernels launched through *launched.
bool sp_can_split(const DScene& sc, uint32_t material_set);
cudaError_t sp_launch_split_level(const DScene& sc, const LevelArgs& a, uint32_t kind_mask, int device, cudaStream_t st, int* launched);
cudaError_t sp_launch_resolve(const ResolveArgs& a, cudaStream_t st);
// accum += scratch unless the chunk's stats say a queue overflowed; scratch = 0 either way
cudaError_t sp_launch_fold(float4* accum, float4* scratch, uint32_t n_pix, const DeviceStats* stats, cudaStream_t st);
cudaError_t sp_launch_add(float4* accum, const float4* other, uint32_t n_pix, cudaStream_t st);   // accum.xyz += other.xyz (other may be peer memory)
cudaError_t sp_upload_decode_tables(const float* plain256, const float* linear256);
// sky-box blur of a cross-layout cube map of packed texels (sp_imaging.cu); tmp0 / tmp1: two (3 * (H / 3))^2 canvases
cudaError_t sp_blur_cube_cross(const uint32_t* cross, uint32_t* out, uint32_t* tmp0, uint32_t* tmp1, int H, int W, float blur,
                               cudaStream_t st);
cudaError_t sp_bench_ffma(double* tflops, cudaStream_t st);
cudaError_t sp_bench_copy(double* gbs, cudaStream_t st);
