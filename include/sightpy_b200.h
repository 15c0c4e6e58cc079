/*
 * sightpy_b200.h — C ABI of the B200-native rendering backend for sightpy.
 *
 * The reference (lmondada/Python-Raytracer) has no FFI: its hot path is the Python call chain
 *     Scene.render            sightpy/scene.py:71-140
 *       -> Camera.get_ray     sightpy/camera.py:51-85
 *       -> get_raycolor       sightpy/ray.py:122-148   (Collider.intersect, Material.get_color ...)
 *       -> accumulate/tonemap sightpy/scene.py:100-140, sightpy/utils/colour_functions.py:4-18
 * This header is the boundary a maintainer binds instead (ctypes stub in INTEGRATION.md): the
 * scene is handed over once as plain-old-data records, then one blocking call renders a frame.
 *
 * Conventions
 *   - every entry point returns 0 on success, non-zero on failure; sp_last_error() returns a
 *     thread-local, NUL-terminated description of the last failure on the calling thread;
 *   - the caller owns every host buffer it passes (C-contiguous); the library owns all device
 *     memory and copies what it needs before returning;
 *   - scene parameters are double precision (they are rounded to float32 on upload; hit points
 *     that feed texel indexing are re-evaluated in double from these values);
 *   - one sp_scene handle is bound to one CUDA device and must not be used from two threads
 *     at once; there is NO CPU fallback: without a usable CUDA device sp_init fails.
 */
#ifndef SIGHTPY_B200_H
#define SIGHTPY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SP_ABI_VERSION 1

/* ---- enums (mirrored in python-raytracer_b200/sightpy/flatten.py) ------------------------- */
enum { SP_COLLIDER_SPHERE = 0, SP_COLLIDER_PLANE = 1, SP_COLLIDER_CUBOID = 2, SP_COLLIDER_TRIANGLE = 3 };
enum { SP_MAT_GLOSSY = 0, SP_MAT_REFRACTIVE = 1, SP_MAT_THINFILM = 2, SP_MAT_DIFFUSE = 3,
       SP_MAT_EMISSIVE = 4, SP_MAT_SKYBOX = 5 };
enum { SP_LIGHT_DIRECTIONAL = 0, SP_LIGHT_POINT = 1 };
enum { SP_DECODE_PLAIN = 0,   /* texel byte b -> b/256                    image_functions.py:7-9   */
       SP_DECODE_LINEAR = 1 };/* texel byte b -> sRGB_to_linear(b/256)     image_functions.py:19-33 */

#define SP_COLLIDER_PAYLOAD 40
#define SP_MAX_DEPTH_LEVELS 64

/* Camera basis as computed by Camera.__init__ (camera.py:8-49). */
typedef struct sp_camera {
    double look_from[3], right[3], up[3], fwd[3];
    double cam_w, cam_h;            /* 2*tan(fov/2), cam_w/aspect                                */
    double lens_radius, focal_distance;
    int32_t width, height;          /* screen_width, screen_height                                */
} sp_camera;

/* One record per Material object (materials/ *.py).  Texture ids index sp_scene_add_texture order;
 * -1 = none / solid colour. */
typedef struct sp_material {
    int32_t kind;                   /* SP_MAT_*                                                   */
    int32_t medium;                 /* Refractive: row of the media table; otherwise -1           */
    int32_t normalmap_tex;          /* Material.normalmap (material.py:18-36)                     */
    int32_t color_tex;              /* image texture of diff_color / color; SkyBox: environment   */
    int32_t aux_tex0;               /* ThinFilm: reflectance LUT; SkyBox: lightmap                */
    int32_t aux_tex1;               /* ThinFilm: thickness noise                                  */
    int32_t diffuse_rays;           /* Diffuse (diffuse.py:13)                                    */
    int32_t max_diffuse_reflections;
    int32_t index_h, index_w;       /* SkyBox: shape used to index the lightmap (skybox.py:74-81) */
    double normalmap_repeat, color_repeat;
    double color[3];                /* solid colour                                               */
    double n_re[3], n_im[3];        /* complex index of refraction (Glossy, Refractive)           */
    double roughness, spec_coeff, diff_coeff;         /* Glossy                                   */
    double thickness, noise_factor;                   /* ThinFilmInterference                     */
    double ambient_weight;                            /* Diffuse: weight of the cosine pdf        */
    double light_intensity;                           /* SkyBox                                   */
} sp_material;

/* One record per Primitive (geometry/primitive.py:6-14). */
typedef struct sp_primitive {
    int32_t material, max_ray_depth, shadow, mc;
    int32_t uv_cross_layout;        /* Cuboid/SkyBox: uv = (u/4, v/3)  (cuboid.py:29-32)          */
    int32_t _pad;
    double center[3];
    double bounded_sphere_radius;   /* used by spherical_caps_pdf (random.py:112-127)             */
} sp_primitive;

/* Tagged collider record; ORDER == scene.collider_list (the index is the "hit id").
 * Payload slots p[]:
 *   sphere   (sphere.py:21-24)    center 0-2, radius 3
 *   plane    (plane.py:39-55)     center 0-2, u_axis 3-5, v_axis 6-8, normal 9-11, w 12, h 13,
 *                                 uv_shift 14-15, inverse_basis_matrix 16-24 (row major)
 *   cuboid   (cuboid.py:60-103)   center 0-2, ax_w 3-5, ax_h 6-8, ax_l 9-11, lb_local 12-14,
 *                                 rt_local 15-17, (width,height,length) 18-20,
 *                                 basis_matrix 21-29, inverse_basis_matrix 30-38 (row major)
 *   triangle (triangle.py:20-35)  p1 0-2, p2 3-5, p3 6-8, normal 9-11, centroid 12-14,
 *                                 n31 15-17, n12 18-20, n23 21-23
 */
typedef struct sp_collider {
    int32_t type;                   /* SP_COLLIDER_*                                              */
    int32_t primitive;
    double p[SP_COLLIDER_PAYLOAD];
} sp_collider;

typedef struct sp_light {           /* lights.py:25-52 */
    int32_t kind, _pad;
    double vec[3];                  /* directional: unit vector towards the light; point: position */
    double color[3];
} sp_light;

/* Counters of one sp_render / sp_trace call. "Rays" follows the reference's definition: one per
 * element of every get_raycolor call (primary + secondary); shadow rays are counted apart. */
typedef struct sp_stats {
    uint64_t rays_total;
    uint64_t shadow_rays;
    uint64_t rays_per_depth[SP_MAX_DEPTH_LEVELS];
    uint64_t kernel_launches;       /* launches of this library's own kernels                     */
    uint64_t chunks;                /* wavefront chunks the frame was split into                  */
    double   device_ms;             /* CUDA-event time of the device work                         */
    double   level_kernel_ms;       /* of which: the fused trace+shade wavefront kernel           */
    uint64_t level_kernel_launches;
    uint64_t queue_bytes;           /* ray-record bytes written + read (algorithmic HBM traffic)  */
    double   level_ms[SP_MAX_DEPTH_LEVELS];   /* level kernel time per recursion depth            */
    uint64_t peak_ray_records;      /* largest per-level queue occupancy seen (records)           */
    uint64_t peak_fan_records;
    uint64_t warp_kernel_launches;  /* of level_kernel_launches: those that ran the warp-autonomous sp_warp_kernel */
    uint64_t chunk_retries;         /* chunks rendered again at half the size after a queue overflow  */
} sp_stats;

typedef struct sp_scene sp_scene;

/* ---- lifetime ------------------------------------------------------------------------------ */
int  sp_abi_version(void);
/* sizeof() of the ABI structs, so that a binding can verify its own layout. */
int  sp_abi_sizes(int32_t out[6]);  /* camera, material, primitive, collider, light, stats */
int  sp_init(int device);           /* bind the calling process to a CUDA device (idempotent)     */
/* Bind the process to several GPUs of one node.  device_ids[0] becomes the default device (new scenes; it gathers
 * and resolves the frames of sp_render_group) and is given peer access to the others. */
int  sp_init_devices(int n, const int* device_ids);
int  sp_default_device(void);       /* -1 before sp_init                                           */
int  sp_device_count(void);
const char* sp_last_error(void);
void sp_shutdown(void);
/* Return idle pooled device memory (wavefront queues of destroyed or re-committed scenes, unreferenced cached
 * textures) to the driver.  On its own the library keeps at most 64 GB of idle buffers per device. */
void sp_trim(void);

int  sp_scene_create(sp_scene** out);               /* on the default device                       */
int  sp_scene_create_on(sp_scene** out, int device);  /* on one of the devices of sp_init_devices    */
void sp_scene_destroy(sp_scene*);

/* ---- scene description (Scene.__init__/add/add_*Light/add_Background, scene.py:29-69) ---------- */
int  sp_scene_set_globals(sp_scene*, const double ambient[3],
                          const double* media_re, const double* media_im, int n_media);
int  sp_scene_set_camera(sp_scene*, const sp_camera*);
int  sp_scene_add_texture(sp_scene*, const uint8_t* rgb_hw3, int H, int W, int decode, int* tex_id);
/* Same, for an image the caller names with a stable non-zero key (same key == same bytes): its texels
 * stay resident on the device and are shared by every later scene of the process that uses the key, so
 * re-describing a scene per animation frame (animation.py:27-31) uploads each image once. */
int  sp_scene_add_texture_keyed(sp_scene*, uint64_t key, const uint8_t* rgb_hw3, int H, int W, int decode,
                                int* tex_id);
/* Same for a cross-layout cube map (3 x 4 blocks of square faces) that the scene wants blurred
 * (Scene.add_Background(..., blur=...), skybox.py:46-49): the library blurs the texels on the device exactly as
 * blur_skybox does with Pillow on the host (blur_background.py:17-132: per face a canvas with its four rotated
 * neighbours, bytes re-quantised, three fixed-point box passes per axis), byte for byte. */
int  sp_scene_add_texture_blurred(sp_scene*, uint64_t key, const uint8_t* rgb_hw3, int H, int W, int decode,
                                  double cube_blur, int* tex_id);
/* Texels of texture tex_id of a committed scene as the device holds them (after any blur), H*W*3 bytes. */
int  sp_scene_read_texture(sp_scene*, int tex_id, uint8_t* out_rgb_hw3);
int  sp_scene_set_materials(sp_scene*, const sp_material*, int n);
int  sp_scene_set_primitives(sp_scene*, const sp_primitive*, int n);
int  sp_scene_set_colliders(sp_scene*, const sp_collider*, int n);
int  sp_scene_set_lights(sp_scene*, const sp_light*, int n);
int  sp_scene_set_importance(sp_scene*, const int32_t* primitive_ids, int n);
int  sp_scene_set_shadow_colliders(sp_scene*, const int32_t* collider_ids, int n);
int  sp_scene_commit(sp_scene*);    /* validate + upload; must precede any render call            */
/* Re-describing a committed scene (animation.py:27-31 calls update_scene + render per frame): any sp_scene_set_* may
 * be called again, followed by sp_scene_commit.  Device buffers (wavefront queues, frame, keyed textures) are
 * recycled, and while the scene keeps its shape (frame size, table sizes, material kinds, collider types) the
 * queue-occupancy estimates of earlier frames are kept, so no probe chunk is rendered again.
 * sp_scene_clear_textures empties the texture list before it is described anew (keyed texels stay resident).
 * sp_scene_update_camera moves the camera of a committed scene without a commit (same frame size only). */
int  sp_scene_clear_textures(sp_scene*);
int  sp_scene_update_camera(sp_scene*, const sp_camera*);

/* ---- rendering ----------------------------------------------------------------------------------
 * sp_render == Scene.render (scene.py:71-140): spp jittered samples per pixel, average, sRGB
 * tonemap, truncation to uint8.  out_linear_rgb (nullable) receives the averaged linear radiance
 * as 3 planes of H*W floats (the layout of the reference's colour.to_array()); out_srgb8
 * (nullable) receives H*W*3 interleaved bytes. */
int  sp_render(sp_scene*, int spp, uint64_t seed, float* out_linear_rgb, uint8_t* out_srgb8,
               sp_stats* stats);

/* Sharded form used for multi-GPU rendering: accumulate samples [sample_begin, sample_end) of
 * every pixel into the scene's device accumulation buffer (float4 per pixel, xyz = sum of
 * radiance).  clear != 0 zeroes the buffer first.  The work is enqueued on the scene's stream and
 * the call returns after it completed. */
int  sp_render_samples(sp_scene*, int sample_begin, int sample_end, uint64_t seed, int clear,
                       sp_stats* stats);
/* Same for the pixels [pix_begin, pix_end) only (row-major pixel index): pixel-band sharding, used when a
 * frame has fewer samples than there are GPUs. */
int  sp_render_region(sp_scene*, int64_t pix_begin, int64_t pix_end, int sample_begin, int sample_end,
                      uint64_t seed, int clear, sp_stats* stats);
/* Same for a list of square tiles (tile_size a power of two; tile ids are row-major over the frame's grid of
 * ceil(W / tile_size) x ceil(H / tile_size) tiles): interleaved-tile sharding — rank r of R renders tiles r, r + R, ...
 * — which balances scenes whose cost varies across the frame.  Replaces the per-batch fan-out of the reference's
 * process pool (scene.py:78-116). */
int  sp_render_tiles(sp_scene*, const int32_t* tile_ids, int n_tiles, int tile_size, int sample_begin, int sample_end,
                     uint64_t seed, int clear, sp_stats* stats);
void*    sp_accum_device_ptr(sp_scene*);      /* float4[H*W] on the scene's device                 */
uint64_t sp_accum_bytes(sp_scene*);
/* Resolve the accumulation buffer: divide by spp_total, tonemap (on the device), then copy to
 * whichever host buffers are non-NULL. */
int  sp_resolve(sp_scene*, int spp_total, float* out_linear_rgb, uint8_t* out_srgb8);
/* Enqueue this scene's work on a caller-owned CUDA stream (a cudaStream_t, e.g. the stream an NCCL
 * reduce of sp_accum_device_ptr is ordered on); use_it == 0 returns to the library's own stream. */
int  sp_scene_set_stream(sp_scene*, void* cuda_stream, int use_it);

/* One frame on several GPUs without a launcher — what the reference's render() does with its process pool
 * (scene.py:98-116).  scenes[0..n) are committed replicas of one scene on n different devices; one host thread per
 * device renders its shard (shard_mode 0: contiguous sample ranges; 1, and whenever spp < n: interleaved 64x64
 * tiles), then scenes[0]'s device adds the other frames to its own — read over NVLink through peer access — and
 * resolves as sp_render does.  stats are summed over the devices (times: the slowest). */
int  sp_render_group(sp_scene** scenes, int n, int spp, uint64_t seed, int shard_mode,
                     float* out_linear_rgb, uint8_t* out_srgb8, sp_stats* stats);

/* sp_trace == get_raycolor(Ray(o, d, depth 0, scene.n), scene) (ray.py:122-148) on caller rays:
 * n rays, origins/directions as n x 3 interleaved floats.  Outputs (each nullable): linear
 * radiance n x 3, index into collider_list of the nearest hit (-1 = none), hit distance. */
int  sp_trace(sp_scene*, const float* origins, const float* dirs, int n, uint64_t seed,
              float* out_rgb, int32_t* out_hit_id, float* out_t, sp_stats* stats);

/* Primary rays of one sample as Camera.get_ray would build them (camera.py:51-85), n = W*H,
 * row-major pixels, interleaved xyz. */
int  sp_camera_rays(sp_scene*, int sample, uint64_t seed, float* out_origins, float* out_dirs);

/* Nearest-hit distance of one jittered primary ray per pixel == ray.get_distances before its
 * clip/normalise step (ray.py:151-163); misses are +inf. */
int  sp_distances(sp_scene*, uint64_t seed, float* out_t);

/* Frame-level debug outputs of one jittered primary ray per pixel (the fields of the reference's Hit record,
 * ray.py:97-119), each nullable: index into collider_list of the nearest hit (-1 = none), hit distance (+inf = none),
 * and the collider normal at the hit point turned towards the ray (collider.get_Normal x orientation; 0 = none),
 * H*W x 3 interleaved. */
int  sp_aovs(sp_scene*, int sample, uint64_t seed, int32_t* out_hit_id, float* out_t, float* out_normal);

/* ---- tuning / measurement ---------------------------------------------------------------------- */
/* options: "ray_queue_capacity", "fan_queue_capacity" (records of 48 B; 0 = auto: 24 records per primary of a chunk,
 *        i.e. 54 GB of queues for a full 8 Mi-primary chunk of a scene with two fan classes, proportionally less for
 *        smaller frames), "chunk_primaries" (0 = auto: up to 8 Mi, more for scenes that queue little; a chunk that
 *        overflows a queue is rendered again at half the size unless "fixed_chunks" = 1),
 * "max_levels" (debugging: trace only the first k recursion depths, 0 = all),
 * "bvh" (1 = scenes with >= 64 colliders put their small colliders into a bounding-volume hierarchy, the
 *        default; 0 = every ray tests every collider),
 * "pretrace" (1 = scenes behind a BVH find the nearest hits of every level with sp_trace_kernel ahead of the level
 *        launch, the default; 0 = inside sp_level_kernel; same hits),
 * "warp_kernel" (1 = queue-fed levels of small untextured Diffuse / Refractive / Emissive scenes run the
 *        warp-autonomous sp_warp_kernel, the default; 0 = the CTA-cooperative sp_level_kernel everywhere; same rays,
 *        same results),
 * "split_kernels" (Whitted scenes — textures, Glossy, Refractive, ThinFilm, sky boxes; no Diffuse, no BVH: a level as a
 *        hit kernel plus one shade kernel per material kind instead of the fused level kernel; 0 = never, 1 = level 0
 *        of launches of 256 Ki primaries or more, the default, 2 = every level; same rays, same results) */
int  sp_set_option(sp_scene*, const char* name, int64_t value);
/* Roofline denominators measured on the bound device: dependent-free FFMA chains (TFLOP/s, 2 flop
 * per FFMA) and a float4 copy (GB/s, read + write bytes). */
int  sp_measure_peaks(double* fp32_tflops, double* copy_gbs);

#ifdef __cplusplus
}
#endif
#endif /* SIGHTPY_B200_H */
