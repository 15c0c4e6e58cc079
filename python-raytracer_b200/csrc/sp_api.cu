// C ABI of the sightpy B200 backend (include/sightpy_b200.h): scene assembly, device residency,
// the chunked wavefront driver and frame resolve.  Host-only code; the kernels live in
// sp_kernels.cu.  Replaces the body of Scene.render (sightpy/scene.py:71-140) and the
// multiprocessing fan-out underneath it.
#include <algorithm>
#include <deque>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "sp_launch.h"

// ---- error handling -------------------------------------------------------------------------------
static thread_local std::string g_error;
static int g_device = -1;          // default device of new scenes (sp_init / first of sp_init_devices)

static int fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    return 1;
}
#define CUDA_TRY(expr)                                                                            \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) return fail("%s: %s", #expr, cudaGetErrorString(e__));            \
    } while (0)

// ---- per-device state ---------------------------------------------------------------------------------------------
// A process may drive several GPUs (sp_init_devices + sp_render_group: one host thread per device).  Everything the
// scenes of a device share lives in that device's context; the calling thread names the device it works on in
// t_device (set by the SP_ENTER guard of every entry point, which also makes it the current CUDA device and takes the
// device's lock, so scenes of different devices render concurrently while calls on one device serialise).
//
// Device allocations are recycled through a per-device free list keyed by size instead of going back to the driver when
// a scene is destroyed: Scene.render after an edit re-commits with the same buffer sizes, and cudaMalloc / cudaFree of
// the multi-GB wavefront queues (plus the device-wide synchronisation every cudaFree implies) would otherwise cost tens
// of milliseconds to seconds per frame.  When a scene is destroyed, idle buffers beyond kPoolIdleLimit are returned
// to the driver (largest first); sp_trim / sp_shutdown return everything.
#include <map>
#include <mutex>
#include <thread>
#define SP_MAX_DEVICES 16
struct CachedTexture { uint64_t key; int H, W, decode; uint32_t* d; size_t bytes; int refs; uint64_t last_use; };
struct DeviceCtx {
    std::recursive_mutex lock;
    bool ready = false;
    std::multimap<size_t, void*> pool;
    size_t pool_bytes = 0;
    // Device-resident textures shared between scenes.  A caller that re-describes a scene every frame
    // (animation.py: update_scene + render) names each image with a stable non-zero key; the packed texels
    // of a key are uploaded once and reused by every later scene of the process (SURVEY §8f "GPU-resident
    // scene reuse": example1's 4096x3072 sky box costs 28 ms to pack and upload, the frame itself 0.6 ms).
    std::vector<CachedTexture> tex_cache;
    uint64_t tex_clock = 0;
    std::vector<cudaStream_t> stream_pool;
    std::vector<cudaEvent_t> event_pool;
};
static DeviceCtx g_ctx[SP_MAX_DEVICES];
static thread_local int t_device = -1;
static DeviceCtx& ctx() { return g_ctx[t_device >= 0 ? t_device : 0]; }
static const size_t kPoolIdleLimit = (size_t)64 << 30;         // idle pooled bytes kept per device (one full-size queue set)
static const size_t kTexCacheLimit = (size_t)8 << 30;          // bytes kept alive for idle (unreferenced) textures

struct DeviceGuard {                     // SP_ENTER(device): this thread works on `device` for the rest of the scope
    int prev;
    std::unique_lock<std::recursive_mutex> lk;
    explicit DeviceGuard(int device) : prev(t_device), lk(g_ctx[device >= 0 && device < SP_MAX_DEVICES ? device : 0].lock) {
        t_device = device >= 0 && device < SP_MAX_DEVICES ? device : 0;
        cudaSetDevice(t_device);
    }
    ~DeviceGuard() { t_device = prev; if (prev >= 0) cudaSetDevice(prev); }
};
#define SP_ENTER(device) DeviceGuard guard__(device)

static void pool_flush() {
    DeviceCtx& c = ctx();
    for (auto& b : c.pool) cudaFree(b.second);
    c.pool.clear();
    c.pool_bytes = 0;
}

static void pool_trim(size_t keep) {     // return the largest idle buffers to the driver until at most `keep` bytes idle
    DeviceCtx& c = ctx();
    while (c.pool_bytes > keep && !c.pool.empty()) {
        auto it = std::prev(c.pool.end());
        cudaFree(it->second);
        c.pool_bytes -= it->first;
        c.pool.erase(it);
    }
}

static cudaError_t pool_get(void** out, size_t bytes) {
    DeviceCtx& c = ctx();
    auto it = c.pool.find(bytes);
    if (it != c.pool.end()) {
        *out = it->second;
        c.pool_bytes -= it->first;
        c.pool.erase(it);
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e != cudaSuccess && !c.pool.empty()) {          // make room and retry once
        cudaGetLastError();
        pool_flush();
        e = cudaMalloc(out, bytes);
    }
    return e;
}

template <typename T> struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    int dev = -1;                        // the device whose pool the buffer came from
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        dev = t_device >= 0 ? t_device : 0;
        return pool_get(reinterpret_cast<void**>(&p), count * sizeof(T));
    }
    cudaError_t upload(const std::vector<T>& h) {
        cudaError_t e = alloc(h.size());
        if (e != cudaSuccess || h.empty()) return e;
        return cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    }
    void release() {
        if (p) {
            DeviceCtx& c = g_ctx[dev >= 0 ? dev : 0];
            c.pool.emplace(n * sizeof(T), p);
            c.pool_bytes += n * sizeof(T);
        }
        p = nullptr; n = 0;
    }
};

struct HostTexture { int H, W, decode; uint64_t key; std::vector<uint32_t> texels; double cube_blur = 0.0; };

static int tex_cache_find(uint64_t key, int H, int W, int decode) {
    auto& tc = ctx().tex_cache;
    for (size_t i = 0; i < tc.size(); ++i)
        if (tc[i].key == key && tc[i].H == H && tc[i].W == W && tc[i].decode == decode)
            return (int)i;
    return -1;
}

static void tex_cache_trim() {
    auto& tc = ctx().tex_cache;
    size_t idle = 0;
    for (auto& t : tc) if (t.refs == 0) idle += t.bytes;
    while (idle > kTexCacheLimit) {
        int victim = -1;
        for (size_t i = 0; i < tc.size(); ++i)
            if (tc[i].refs == 0 && (victim < 0 || tc[i].last_use < tc[(size_t)victim].last_use)) victim = (int)i;
        if (victim < 0) break;
        idle -= tc[(size_t)victim].bytes;
        cudaFree(tc[(size_t)victim].d);
        tc.erase(tc.begin() + victim);
    }
}

struct QueueSet {
    DevBuf<float4> q0, q1, q2;
    size_t n = 0;
    RayQueue view(uint32_t cap) { RayQueue q; q.q0 = q0.p; q.q1 = q1.p; q.q2 = q2.p; q.capacity = cap; return q; }
    cudaError_t alloc(size_t count) {
        cudaError_t e;
        n = count;
        if ((e = q0.alloc(count)) != cudaSuccess) return e;
        if ((e = q1.alloc(count)) != cudaSuccess) return e;
        return q2.alloc(count);
    }
    void release() { q0.release(); q1.release(); q2.release(); n = 0; }
};

// streams and events are recycled the same way (per device)
struct ScopedEvent {                     // a pooled event for the duration of one call
    cudaEvent_t e = nullptr;
    int dev;
    ScopedEvent() : dev(t_device >= 0 ? t_device : 0) {
        auto& pool = g_ctx[dev].event_pool;
        if (!pool.empty()) { e = pool.back(); pool.pop_back(); }
        else if (cudaEventCreate(&e) != cudaSuccess) e = nullptr;
    }
    ~ScopedEvent() { if (e) g_ctx[dev].event_pool.push_back(e); }
};

#define SPS_KINDS_HOST 6        // shading bins of the split kernels (sp_split_kernels.cuh: SPS_KINDS)
struct ChunkJob {
    int source, run;
    uint32_t pix_begin, n_pix, sample_begin, n_items, user_base;
    const uint32_t* tiles; uint32_t tile_shift, tiles_x;
    const float* user_o; const float* user_d;
    float4* accum; int32_t* out_hit; float* out_t; float* out_o; float* out_d; float* out_n;
};


// What one wavefront chunk in flight owns.  Two slots: the host enqueues chunk k + 1 (all its levels, the copy of its
// counters into pinned memory, its fold into the frame) before it waits for chunk k, so that the GPU never idles
// between the chunks of a frame on a host round trip.
struct ChunkSlot {
    DevBuf<uint32_t> counts;            // per level: queue counts + work counters (SP_COUNTS_PER_LEVEL words)
    DevBuf<DeviceStats> d_stats;
    DevBuf<float4> scratch;             // the chunk's radiance until it is known to be complete (frame-sized)
    std::vector<cudaEvent_t> events;    // level boundaries
    cudaEvent_t done = nullptr;         // everything of the chunk, copies included
    uint32_t* h_counts = nullptr;       // pinned host copies
    DeviceStats* h_stats = nullptr;
    // the chunk in flight
    bool busy = false;
    ChunkJob job{};
    int n_levels = 0, split_mode = 0, split_launches = 0;
    bool warp = false, pretrace = false, defer_shadows = false, split = false, folded = false;
    uint32_t region_pix = 0, region_ns = 0;       // what to render again if it overflows (offset / samples of the job)
};

struct sp_scene {
    int device = 0;                     // the CUDA device this scene lives on
    // ---- host description ---------------------------------------------------------------------
    double ambient[3] = {0, 0, 0};
    std::vector<double> media_re, media_im;
    sp_camera cam{};
    bool has_camera = false, committed = false;
    bool user_stream_set = false;
    cudaStream_t user_stream = nullptr;
    std::vector<HostTexture> textures;
    std::vector<sp_material> mats;
    std::vector<sp_primitive> prims;
    std::vector<sp_collider> cols;
    std::vector<sp_light> lights;
    std::vector<int32_t> importance, shadow_ids;
    // ---- options ------------------------------------------------------------------------------------
    int64_t opt_ray_cap = 0, opt_fan_cap = 0, opt_chunk = 0, opt_max_levels = 0, opt_bvh = 1, opt_warp = 1;
    int64_t opt_split = 1;                       // Whitted scenes: hit kernel + per-material shade kernels (sp_split_kernels.cuh)
    uint32_t kind_mask = 0;                      // material kinds present (bit SP_MAT_*)
    uint64_t call_primaries = 0;                 // primaries of the render call in progress (sizes per-item buffers once)
    int64_t opt_pretrace = 1;                    // BVH scenes: sp_trace_kernel ahead of every level launch
    int64_t opt_chunk_fixed = 0;                 // 1: a chunk that overflows is an error instead of being retried smaller
    int64_t chunk_limit = 0;                     // learnt from overflows: no chunk larger than this
    uint64_t shape_sig = 0;                      // what the occupancy estimates below were measured on (shape_signature)
    // ---- device residency -----------------------------------------------------------------------------
    DScene d{};
    int n_levels = 1;
    cudaStream_t stream = nullptr;      // where all work of this scene is enqueued
    cudaStream_t own_stream = nullptr;  // the library's default stream for it
    DevBuf<float> d_lin;                // resolve outputs (device copies, reused across frames)
    DevBuf<uint8_t> d_u8;
    ChunkSlot slot[2];
    DevBuf<float4> geom_all, geom_shadow, accum;
    DevBuf<uint32_t> d_tiles;                    // tile list of the last sp_render_tiles call
    DevBuf<float4> d_shq;                        // BVH scenes: shadow-ray requests of the level that just ran (sp_shadow_kernel)
    DevBuf<uint32_t> d_shq_count;                // per level: requests queued, work counter
    DevBuf<float2> d_hits;                       // BVH scenes: per-item nearest hits of the level about to run (sp_trace_kernel)
    DevBuf<uint32_t> d_klist, d_kcount;          // Whitted scenes: per-kind item lists of the level being shaded and their counts
    DevBuf<int> off_all, off_shadow;
    DevBuf<int2> slot_shadow;
    DevBuf<DCollider> d_cols;
    DevBuf<float4> bvh_nodes, bvh_data;
    DevBuf<int4> bvh_items;
    DevBuf<DColInfo> d_colinfo;
    DevBuf<float4> d_collite;
    DevBuf<double> d_cols_d;
    DevBuf<DPrimitive> d_prims;
    DevBuf<DMaterial> d_mats;
    DevBuf<DTexture> d_texdesc;
    std::vector<DevBuf<uint32_t>> d_texels;
    std::vector<uint64_t> cached_tex_keys;       // keys of the shared textures this scene holds a reference to
    DevBuf<DMedium> d_media;
    QueueSet ray_q[2], fan_q[2];
    uint32_t ray_cap = 0, fan_cap = 0;
    uint32_t chunk_primaries = 0;
    double use_ray = 0.0, use_fan = 0.0;         // peak records per primary seen so far
    int grid0 = 0, grid_q = 0;                   // CTAs of level-0 / queue-fed launches of sp_level_kernel
    uint32_t material_set = 0;                   // compiled kernel variant (sp_pick_material_set)

    ~sp_scene() {
        release_device(); release_textures();
        for (auto& sl : slot) {
            if (sl.h_counts) cudaFreeHost(sl.h_counts);
            if (sl.h_stats) cudaFreeHost(sl.h_stats);
            sl.h_counts = nullptr; sl.h_stats = nullptr;
        }
        pool_trim(kPoolIdleLimit);           // what a destroyed scene leaves idle beyond the limit goes back to the driver
    }
    void release_textures() {                    // a scene holds one reference per shared texture for its whole life
        for (uint64_t k : cached_tex_keys)
            for (auto& t : g_ctx[device].tex_cache) if (t.key == k && t.refs > 0) { t.refs--; break; }
        cached_tex_keys.clear();
        tex_cache_trim();
    }
    bool holds_texture(uint64_t key) const {
        return std::find(cached_tex_keys.begin(), cached_tex_keys.end(), key) != cached_tex_keys.end();
    }
    void release_device() {
        for (auto& b : d_texels) b.release();
        d_texels.clear();

        geom_all.release(); geom_shadow.release(); accum.release(); d_tiles.release(); d_hits.release(); d_klist.release(); d_kcount.release(); d_shq.release(); d_shq_count.release(); off_all.release(); off_shadow.release();
        slot_shadow.release(); d_cols.release(); d_colinfo.release(); d_collite.release(); bvh_nodes.release(); bvh_data.release(); bvh_items.release(); d_cols_d.release(); d_prims.release();
        d_mats.release(); d_texdesc.release(); d_media.release();
        for (int i = 0; i < 2; ++i) { ray_q[i].release(); fan_q[i].release(); }
        for (auto& sl : slot) {
            sl.counts.release(); sl.d_stats.release(); sl.scratch.release();
            for (auto e : sl.events) g_ctx[device].event_pool.push_back(e);
            sl.events.clear();
            if (sl.done) { g_ctx[device].event_pool.push_back(sl.done); sl.done = nullptr; }
            sl.busy = false;                 // (the pinned host copies are fixed-size: kept until the scene is destroyed)
        }
        if (own_stream) { cudaStreamSynchronize(own_stream); g_ctx[device].stream_pool.push_back(own_stream); }
        own_stream = nullptr;
        stream = nullptr;
        d_lin.release(); d_u8.release();
        ray_cap = fan_cap = 0;
    }
};

static float3 f3(const double* v) { return make_float3((float)v[0], (float)v[1], (float)v[2]); }

// ---- geometry stream construction (layout documented in sp_types.cuh) ------------------------------
// stream type of a collider: bounded planes whose normal and edge axes are coordinate axes go to the
// cheap axis-aligned sections
static int unit_axis(const double* v) {          // index of the axis v is +-1 along, -1 if none
    int axis = -1;
    for (int k = 0; k < 3; ++k) {
        if (v[k] == 0.0) continue;
        if (std::fabs(v[k]) != 1.0 || axis >= 0) return -1;
        axis = k;
    }
    return axis;
}

static int stream_type(const sp_collider& c) {
    if (c.type != SP_COLLIDER_PLANE) return c.type;          // SP_ST_* == SP_COLLIDER_* for the general shapes
    const int an = unit_axis(c.p + 9), au = unit_axis(c.p + 3), av = unit_axis(c.p + 6);
    if (an < 0 || au < 0 || av < 0 || an == au || an == av || au == av) return SP_ST_PLANE;
    return SP_ST_AAX + an;
}

static int type_vec4(int st) {
    switch (st) {
    case SP_ST_SPHERE: return SP_V4_SPHERE;
    case SP_ST_PLANE: return SP_V4_PLANE;
    case SP_ST_CUBOID: return SP_V4_CUBOID;
    case SP_ST_TRI: return SP_V4_TRIANGLE;
    default: return SP_V4_AARECT;
    }
}

static void pack_collider(const sp_collider& c, int st, float* out) {
    const double* p = c.p;
    auto put3 = [&](int at, double x, double y, double z) { out[at] = (float)x; out[at + 1] = (float)y; out[at + 2] = (float)z; };
    switch (st) {
    case SP_ST_SPHERE:
        put3(0, p[0], p[1], p[2]); out[3] = (float)(p[3] * p[3]);
        break;
    case SP_ST_PLANE:
        put3(0, p[9], p[10], p[11]);  out[3] = (float)p[12];        // N, w
        put3(4, p[0], p[1], p[2]);    out[7] = (float)p[13];        // C, h
        put3(8, p[3], p[4], p[5]);    out[11] = 0.f;                // U
        put3(12, p[6], p[7], p[8]);   out[15] = 0.f;                // V
        break;
    case SP_ST_CUBOID: {
        const double* B = p + 21;
        double bc[3], lo[3], hi[3];
        for (int r = 0; r < 3; ++r) {
            bc[r] = B[3 * r] * p[0] + B[3 * r + 1] * p[1] + B[3 * r + 2] * p[2];
            lo[r] = p[12 + r] - bc[r];
            hi[r] = p[15 + r] - bc[r];
        }
        put3(0, B[0], B[1], B[2]);  out[3] = (float)lo[0];
        put3(4, B[3], B[4], B[5]);  out[7] = (float)lo[1];
        put3(8, B[6], B[7], B[8]);  out[11] = (float)lo[2];
        put3(12, p[0], p[1], p[2]); out[15] = (float)hi[0];
        out[16] = (float)hi[1]; out[17] = (float)hi[2]; out[18] = 0.f; out[19] = 0.f;
        break;
    }
    case SP_ST_TRI: {   // rows of M = [e1 e2 N]^-1 and t = -M p1 (sp_types.cuh)
        const double e1[3] = {p[3] - p[0], p[4] - p[1], p[5] - p[2]}, e2[3] = {p[6] - p[0], p[7] - p[1], p[8] - p[2]};
        const double* N = p + 9;
        auto cross3 = [](const double* a, const double* b, double* o) {
            o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
        };
        double c0[3], c1[3], c2[3];                     // rows of the inverse: (e2 x N, N x e1, e1 x e2) / det
        cross3(e2, N, c0); cross3(N, e1, c1); cross3(e1, e2, c2);
        const double det = e1[0] * c0[0] + e1[1] * c0[1] + e1[2] * c0[2];
        const double inv = det != 0.0 ? 1.0 / det : 0.0;  // degenerate triangles can never be hit (u, v = 0 + t*0 ... t0 = 0)
        const double* rows[3] = {c0, c1, c2};
        for (int r = 0; r < 3; ++r) {
            double row[3] = {rows[r][0] * inv, rows[r][1] * inv, rows[r][2] * inv};
            if (r == 2) { row[0] = N[0]; row[1] = N[1]; row[2] = N[2]; }      // exactly the reference normal
            put3(4 * r, row[0], row[1], row[2]);
            out[4 * r + 3] = (float)(-(row[0] * p[0] + row[1] * p[1] + row[2] * p[2]));
        }
        break;
    }
    default: {        // axis-aligned rectangle: (C, sign of N) (half extents along the in-plane axes, ascending)
        const int an = st - SP_ST_AAX, au = unit_axis(p + 3);
        const int b = an == 0 ? 1 : 0;                   // lower in-plane axis
        put3(0, p[0], p[1], p[2]); out[3] = (float)p[9 + an];
        out[4] = (float)(au == b ? p[12] : p[13]);       // w belongs to u_axis, h to v_axis
        out[5] = (float)(au == b ? p[13] : p[12]);
        out[6] = 0.f; out[7] = 0.f;
        break;
    }
    }
}

struct BuiltStream {
    std::vector<float4> data;
    std::vector<int> chunk_off;
    std::vector<int2> slot;          // per collider id of the scene
    int n_items = 0;
};

static BuiltStream build_stream(const std::vector<sp_collider>& cols, std::vector<int32_t> ids) {
    BuiltStream bs;
    bs.slot.assign(cols.size(), make_int2(-1, -1));
    bs.n_items = (int)ids.size();
    std::vector<int> st(cols.size());
    for (size_t i = 0; i < cols.size(); ++i) st[i] = stream_type(cols[i]);
    std::stable_sort(ids.begin(), ids.end(), [&](int a, int b) { return st[a] < st[b]; });
    bs.chunk_off.push_back(0);
    size_t pos = 0;
    while (pos < ids.size()) {
        // greedy fill of one chunk
        int cnt[7] = {0, 0, 0, 0, 0, 0, 0}, vec4 = 4;
        size_t end = pos;
        while (end < ids.size()) {
            const int t = st[ids[end]];
            const int items = (int)(end - pos) + 1;
            const int need = vec4 + type_vec4(t) + (items + 3) / 4;
            if (need > SP_CHUNK_VEC4) break;
            vec4 += type_vec4(t);
            cnt[t]++;
            ++end;
        }
        const int items = (int)(end - pos);
        GeomChunkHeader h{};
        h.n_sphere = cnt[0]; h.n_plane = cnt[1]; h.n_cuboid = cnt[2]; h.n_tri = cnt[3];
        h.n_aax = cnt[4]; h.n_aay = cnt[5]; h.n_aaz = cnt[6];
        h.off_sphere = 4;
        h.off_plane = h.off_sphere + cnt[0] * SP_V4_SPHERE;
        h.off_cuboid = h.off_plane + cnt[1] * SP_V4_PLANE;
        h.off_tri = h.off_cuboid + cnt[2] * SP_V4_CUBOID;
        h.off_aa = h.off_tri + cnt[3] * SP_V4_TRIANGLE;
        h.off_ids = h.off_aa + (cnt[4] + cnt[5] + cnt[6]) * SP_V4_AARECT;
        h.n_vec4 = h.off_ids + (items + 3) / 4;
        std::vector<float4> chunk((size_t)h.n_vec4, make_float4(0, 0, 0, 0));
        memcpy(chunk.data(), &h, sizeof h);
        float* fl = reinterpret_cast<float*>(chunk.data());
        int* idp = reinterpret_cast<int*>(chunk.data() + h.off_ids);
        int at = 4 * 4, local[7] = {0, 0, 0, 0, 0, 0, 0};
        const int chunk_index = (int)bs.chunk_off.size() - 1;
        for (size_t k = pos; k < end; ++k) {
            const sp_collider& c = cols[ids[k]];
            const int t = st[ids[k]];
            pack_collider(c, t, fl + at);
            if (t >= SP_ST_AAX) {
                // third word of the second vector: position in the id array | (normal along +axis) << 31, which
                // sp_intersect_lean turns into tag | orientation with one XOR (sp_geometry.cuh)
                const uint32_t code = (uint32_t)(k - pos) | (fl[at + 3] > 0.f ? 0x80000000u : 0u);
                memcpy(fl + at + 6, &code, sizeof code);
            }
            at += 4 * type_vec4(t);
            idp[k - pos] = ids[k];
            // the three axis-aligned sections share one index space (code 4)
            const int aa_index = t == SP_ST_AAX ? local[4] : t == SP_ST_AAY ? cnt[4] + local[5] : cnt[4] + cnt[5] + local[6];
            bs.slot[ids[k]] = t >= SP_ST_AAX ? make_int2(chunk_index, (4 << 28) | aa_index)
                                             : make_int2(chunk_index, (t << 28) | local[t]);
            local[t]++;
        }
        bs.data.insert(bs.data.end(), chunk.begin(), chunk.end());
        bs.chunk_off.push_back((int)bs.data.size());
        pos = end;
    }
    if (ids.empty()) {                       // an empty stream still has one (empty) chunk
        GeomChunkHeader h{};
        h.off_sphere = h.off_plane = h.off_cuboid = h.off_tri = h.off_aa = h.off_ids = h.n_vec4 = 4;
        std::vector<float4> chunk(4, make_float4(0, 0, 0, 0));
        memcpy(chunk.data(), &h, sizeof h);
        bs.data = chunk;
        bs.chunk_off.push_back(4);
    }
    return bs;
}

// ---- BVH over the small colliders of a large scene (layout and rationale: sp_geometry.cuh) ----------------
struct Aabb { double lo[3], hi[3]; };

static Aabb collider_aabb(const sp_collider& c) {
    const double* p = c.p;
    Aabb b;
    auto from_center = [&](const double* ctr, const double ext[3]) {
        for (int k = 0; k < 3; ++k) { b.lo[k] = ctr[k] - ext[k]; b.hi[k] = ctr[k] + ext[k]; }
    };
    switch (c.type) {
    case SP_COLLIDER_SPHERE: { const double e[3] = {std::fabs(p[3]), std::fabs(p[3]), std::fabs(p[3])}; from_center(p, e); break; }
    case SP_COLLIDER_PLANE: {            // |u_axis . (M - C)| <= w, |v_axis . (M - C)| <= h  (axes as given, plane.py:57-90)
        double e[3];
        const double uu = p[3] * p[3] + p[4] * p[4] + p[5] * p[5], vv = p[6] * p[6] + p[7] * p[7] + p[8] * p[8];
        for (int k = 0; k < 3; ++k)
            e[k] = (uu > 0 ? std::fabs(p[3 + k]) / uu * std::fabs(p[12]) : 0.0) + (vv > 0 ? std::fabs(p[6 + k]) / vv * std::fabs(p[13]) : 0.0);
        from_center(p, e);
        break;
    }
    case SP_COLLIDER_CUBOID: {           // centre +- sum of |axis| * half size
        double e[3];
        for (int k = 0; k < 3; ++k)
            e[k] = 0.5 * (std::fabs(p[3 + k] * p[18]) + std::fabs(p[6 + k] * p[19]) + std::fabs(p[9 + k] * p[20]));
        from_center(p, e);
        break;
    }
    default:                             // triangle
        for (int k = 0; k < 3; ++k) {
            b.lo[k] = std::min(p[k], std::min(p[3 + k], p[6 + k]));
            b.hi[k] = std::max(p[k], std::max(p[3 + k], p[6 + k]));
        }
    }
    for (int k = 0; k < 3; ++k) {        // conservative: beyond anything float rounding of the slab test can cost
        const double pad = 1e-5 * (std::fabs(b.lo[k]) + std::fabs(b.hi[k])) + 1e-5;
        b.lo[k] -= pad; b.hi[k] += pad;
    }
    return b;
}

struct BuiltBvh {
    std::vector<float4> nodes, data;
    std::vector<int4> items;
};

static int bvh_build_rec(BuiltBvh& out, std::vector<int>& order, int first, int count, const std::vector<Aabb>& boxes,
                         const std::vector<int>& item_of, Aabb& bounds, int depth = 0) {
    // returns the child code of this subtree (>= 0 node index, < 0 leaf) and its bounds
    for (int k = 0; k < 3; ++k) { bounds.lo[k] = 1e300; bounds.hi[k] = -1e300; }
    for (int i = first; i < first + count; ++i)
        for (int k = 0; k < 3; ++k) {
            bounds.lo[k] = std::min(bounds.lo[k], boxes[order[i]].lo[k]);
            bounds.hi[k] = std::max(bounds.hi[k], boxes[order[i]].hi[k]);
        }
    static const int leaf_max = [] { const char* e = getenv("SIGHTPY_BVH_LEAF"); const int v = e ? atoi(e) : 2; return v < 1 ? 1 : (v > 8 ? 8 : v); }();
    if (count <= leaf_max) {             // leaf: its items become consecutive in the item array, by type, then by id
        const int at = (int)out.items.size();
        std::sort(order.begin() + first, order.begin() + first + count, [&](int a, int b) {
            const int ta = item_of[a] & 255, tb = item_of[b] & 255;
            return ta != tb ? ta < tb : a < b;
        });
        for (int i = first; i < first + count; ++i) out.items.push_back(make_int4(item_of[order[i]], 0, order[i], 0));
        return ~((at << 3) | (count - 1));
    }
    double clo[3] = {1e300, 1e300, 1e300}, chi[3] = {-1e300, -1e300, -1e300};
    for (int i = first; i < first + count; ++i)
        for (int k = 0; k < 3; ++k) {
            const double c = 0.5 * (boxes[order[i]].lo[k] + boxes[order[i]].hi[k]);
            clo[k] = std::min(clo[k], c); chi[k] = std::max(chi[k], c);
        }
    // Binned surface-area heuristic: 16 bins of box centres per axis, the split that minimises
    // area(left) * n_left + area(right) * n_right.  Falls back to the median along the widest axis when no bin
    // boundary separates the centres, or when the tree gets so deep that only balanced splits keep it within the
    // traversal stack (32 entries).
    auto half_area = [](const Aabb& b) {
        const double dx = std::max(b.hi[0] - b.lo[0], 0.0), dy = std::max(b.hi[1] - b.lo[1], 0.0), dz = std::max(b.hi[2] - b.lo[2], 0.0);
        return dx * dy + dy * dz + dz * dx;
    };
    const int NB = 16;
    int best_axis = -1, best_bin = -1;
    double best_cost = 1e300;
    int log2_count = 0;
    while ((1 << log2_count) < count) ++log2_count;
    if (depth + log2_count < 28) {                           // a median split from here on still fits the 32-entry stack
        for (int k = 0; k < 3; ++k) {
            const double ext = chi[k] - clo[k];
            if (!(ext > 0.0)) continue;
            Aabb bb_[NB]; int cnt[NB];
            for (int b = 0; b < NB; ++b) { cnt[b] = 0; for (int q = 0; q < 3; ++q) { bb_[b].lo[q] = 1e300; bb_[b].hi[q] = -1e300; } }
            for (int i = first; i < first + count; ++i) {
                const Aabb& bx = boxes[order[i]];
                const double c = 0.5 * (bx.lo[k] + bx.hi[k]);
                const int b = std::min(NB - 1, (int)((c - clo[k]) / ext * NB));
                cnt[b]++;
                for (int q = 0; q < 3; ++q) { bb_[b].lo[q] = std::min(bb_[b].lo[q], bx.lo[q]); bb_[b].hi[q] = std::max(bb_[b].hi[q], bx.hi[q]); }
            }
            double right_area[NB]; int right_cnt[NB];
            Aabb acc; for (int q = 0; q < 3; ++q) { acc.lo[q] = 1e300; acc.hi[q] = -1e300; }
            int n = 0;
            for (int b = NB - 1; b > 0; --b) {
                for (int q = 0; q < 3; ++q) { acc.lo[q] = std::min(acc.lo[q], bb_[b].lo[q]); acc.hi[q] = std::max(acc.hi[q], bb_[b].hi[q]); }
                n += cnt[b];
                right_area[b] = n ? half_area(acc) : 0.0; right_cnt[b] = n;
            }
            for (int q = 0; q < 3; ++q) { acc.lo[q] = 1e300; acc.hi[q] = -1e300; }
            n = 0;
            for (int b = 0; b + 1 < NB; ++b) {                 // split after bin b
                for (int q = 0; q < 3; ++q) { acc.lo[q] = std::min(acc.lo[q], bb_[b].lo[q]); acc.hi[q] = std::max(acc.hi[q], bb_[b].hi[q]); }
                n += cnt[b];
                if (n == 0 || right_cnt[b + 1] == 0) continue;
                const double cost = half_area(acc) * n + right_area[b + 1] * right_cnt[b + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = k; best_bin = b; }
            }
        }
    }
    int half;
    if (best_axis >= 0) {
        const int k = best_axis;
        const double ext = chi[k] - clo[k];
        auto bin_of = [&](int id) {
            const double c = 0.5 * (boxes[id].lo[k] + boxes[id].hi[k]);
            return std::min(NB - 1, (int)((c - clo[k]) / ext * NB));
        };
        auto mid = std::partition(order.begin() + first, order.begin() + first + count, [&](int id) { return bin_of(id) <= best_bin; });
        half = (int)(mid - (order.begin() + first));
    } else {
        int axis = 0;
        for (int k = 1; k < 3; ++k) if (chi[k] - clo[k] > chi[axis] - clo[axis]) axis = k;
        half = count / 2;
        std::nth_element(order.begin() + first, order.begin() + first + half, order.begin() + first + count, [&](int a, int b) {
            return boxes[a].lo[axis] + boxes[a].hi[axis] < boxes[b].lo[axis] + boxes[b].hi[axis];
        });
    }
    const int me = (int)out.nodes.size() / 4;
    out.nodes.resize(out.nodes.size() + 4);
    Aabb ba, bb;
    const int ca = bvh_build_rec(out, order, first, half, boxes, item_of, ba, depth + 1);
    const int cb = bvh_build_rec(out, order, first + half, count - half, boxes, item_of, bb, depth + 1);
    out.nodes[4 * me] = make_float4((float)ba.lo[0], (float)ba.lo[1], (float)ba.lo[2], (float)ba.hi[0]);
    out.nodes[4 * me + 1] = make_float4((float)ba.hi[1], (float)ba.hi[2], (float)bb.lo[0], (float)bb.lo[1]);
    out.nodes[4 * me + 2] = make_float4((float)bb.lo[2], (float)bb.hi[0], (float)bb.hi[1], (float)bb.hi[2]);
    float4 kids; memset(&kids, 0, sizeof kids);
    memcpy(&kids.x, &ca, 4); memcpy(&kids.y, &cb, 4);
    out.nodes[4 * me + 3] = kids;
    return me;
}

// Which colliders go into the BVH: finite ones whose box is small against the extent of all collider centres
// (a ground plane or a sky box would sit at the root and be tested by every ray anyway).
static std::vector<char> bvh_membership(const std::vector<sp_collider>& cols, const std::vector<Aabb>& boxes) {
    std::vector<char> in((size_t)cols.size(), 0);
    if ((int)cols.size() < SP_BVH_MIN_COLLIDERS) return in;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    std::vector<double> diag(cols.size());
    for (size_t i = 0; i < cols.size(); ++i) {
        double d2 = 0;
        for (int k = 0; k < 3; ++k) { const double e = boxes[i].hi[k] - boxes[i].lo[k]; d2 += e * e; }
        diag[i] = std::sqrt(d2);
    }
    std::vector<double> sorted(diag);
    std::nth_element(sorted.begin(), sorted.begin() + sorted.size() / 2, sorted.end());
    const double typical = sorted[sorted.size() / 2];            // median collider size
    for (size_t i = 0; i < cols.size(); ++i)
        for (int k = 0; k < 3; ++k) {
            const double c = 0.5 * (boxes[i].lo[k] + boxes[i].hi[k]);
            if (diag[i] <= 64.0 * typical) { lo[k] = std::min(lo[k], c); hi[k] = std::max(hi[k], c); }
        }
    double extent = 0;
    for (int k = 0; k < 3; ++k) extent = std::max(extent, hi[k] - lo[k]);
    int n_in = 0;
    for (size_t i = 0; i < cols.size(); ++i) {
        in[i] = (diag[i] <= 0.25 * extent && std::isfinite(diag[i])) ? 1 : 0;
        n_in += in[i];
    }
    if (n_in < SP_BVH_MIN_COLLIDERS / 2) std::fill(in.begin(), in.end(), 0);
    return in;
}

static BuiltBvh build_bvh(const std::vector<sp_collider>& cols, const std::vector<char>& member, const std::vector<Aabb>& boxes,
                          const std::vector<char>& casts_shadow) {
    BuiltBvh out;
    std::vector<int> order, item_of(cols.size(), 0);
    for (size_t i = 0; i < cols.size(); ++i) {
        if (!member[i]) continue;
        order.push_back((int)i);
        item_of[i] = stream_type(cols[i]) | (casts_shadow[i] ? 256 : 0);
    }
    if (order.empty()) return out;
    Aabb root;
    const int code = bvh_build_rec(out, order, 0, (int)order.size(), boxes, item_of, root);
    if (code < 0) {                     // a single leaf: wrap it in a root whose second child is an empty box
        out.nodes.assign(4, make_float4(0, 0, 0, 0));
        out.nodes[0] = make_float4((float)root.lo[0], (float)root.lo[1], (float)root.lo[2], (float)root.hi[0]);
        out.nodes[1] = make_float4((float)root.hi[1], (float)root.hi[2], 1.f, 1.f);
        out.nodes[2] = make_float4(1.f, -1.f, -1.f, -1.f);
        float4 kids; memset(&kids, 0, sizeof kids);
        const int none = code;          // never visited: its box is inverted
        memcpy(&kids.x, &code, 4); memcpy(&kids.y, &none, 4);
        out.nodes[3] = kids;
    }
    // Leaf records, in item order: one header vector (stream type | casts shadow << 8, collider id, data vectors, -)
    // followed by the collider's packed data, so that a leaf is one contiguous run the traversal walks with a single
    // pointer (no item table between the node and the data).
    std::vector<int> record_at(out.items.size());
    for (size_t k = 0; k < out.items.size(); ++k) {
        int4& it = out.items[k];
        const int st = it.x & 255, nv = type_vec4(st);
        record_at[k] = (int)out.data.size();
        it.y = record_at[k] + 1;
        float4 head; memset(&head, 0, sizeof head);
        memcpy(&head.x, &it.x, 4); memcpy(&head.y, &it.z, 4); memcpy(&head.z, &nv, 4);
        out.data.push_back(head);
        std::vector<float> tmp((size_t)4 * nv, 0.f);
        pack_collider(cols[(size_t)it.z], st, tmp.data());
        for (int v = 0; v < nv; ++v) out.data.push_back(make_float4(tmp[4 * v], tmp[4 * v + 1], tmp[4 * v + 2], tmp[4 * v + 3]));
    }
    // leaf codes name the first item of the leaf: turn them into the offset of its record
    for (size_t nd = 0; nd + 3 < out.nodes.size(); nd += 4)
        for (int k = 0; k < 2; ++k) {
            float* slot = k == 0 ? &out.nodes[nd + 3].x : &out.nodes[nd + 3].y;
            int code; memcpy(&code, slot, 4);
            if (code >= 0) continue;
            const int c = ~code, first = c >> 3, cnt1 = c & 7;
            const int recoded = ~((record_at[(size_t)first] << 3) | cnt1);
            memcpy(slot, &recoded, 4);
        }
    return out;
}

// Texels of a cross-layout cube map that the scene wants blurred (add_Background(..., blur=...), skybox.py:46-49): the
// device copy `d` (H x W packed texels) is replaced by its blurred version (sp_imaging.cu).
static int blur_texture_in_place(uint32_t* d, int H, int W, double blur) {
    const int N = H / 3;
    if (N < 1 || 4 * N > W) return fail("cube-map blur: a %dx%d image is not a 3 x 4 cross of square faces", W, H);
    DevBuf<uint32_t> out, tmp0, tmp1;
    cudaError_t e = out.alloc((size_t)H * W);
    if (e == cudaSuccess) e = tmp0.alloc((size_t)9 * N * N);
    if (e == cudaSuccess) e = tmp1.alloc((size_t)9 * N * N);
    if (e == cudaSuccess) e = sp_blur_cube_cross(d, out.p, tmp0.p, tmp1.p, H, W, (float)blur, nullptr);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d, out.p, (size_t)H * W * sizeof(uint32_t), cudaMemcpyDeviceToDevice, nullptr);
    if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
    out.release(); tmp0.release(); tmp1.release();
    if (e != cudaSuccess) return fail("cube-map blur: %s", cudaGetErrorString(e));
    return 0;
}

// Fingerprint of everything that decides how many records a primary ray puts into the queues *structurally*: frame
// size, table sizes, material kinds / fan sizes / depth limits, collider types.  A scene that is re-described with
// the same shape (an animation frame: moved colliders, another camera pose) keeps the queue-occupancy estimates of
// the previous frames instead of probing again.
static uint64_t shape_signature(const sp_scene* s) {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&h](uint64_t v) { h ^= v; h *= 1099511628211ull; };
    mix((uint64_t)s->cam.width); mix((uint64_t)s->cam.height);
    mix(s->textures.size()); mix(s->lights.size()); mix(s->importance.size()); mix(s->shadow_ids.size());
    for (const auto& t : s->textures) { mix((uint64_t)t.H); mix((uint64_t)t.W); }
    for (const auto& m : s->mats) { mix((uint64_t)m.kind); mix((uint64_t)m.diffuse_rays); mix((uint64_t)m.max_diffuse_reflections); }
    for (const auto& p : s->prims) { mix((uint64_t)p.material); mix((uint64_t)p.max_ray_depth); mix((uint64_t)p.mc); }
    for (const auto& c : s->cols) { mix((uint64_t)c.type); mix((uint64_t)c.primitive); }
    return h ? h : 1;
}

// Launch geometry of the level kernels for this scene.
static int pick_kernels(sp_scene* s) {
    s->grid0 = sp_level_grid(s->device, s->d, s->material_set, true);
    s->grid_q = sp_level_grid(s->device, s->d, s->material_set, false);
    return 0;
}

// =================================================================================================
// lifetime
// =================================================================================================
extern "C" {

int sp_abi_version(void) { return SP_ABI_VERSION; }

int sp_abi_sizes(int32_t out[6]) {
    out[0] = (int32_t)sizeof(sp_camera);   out[1] = (int32_t)sizeof(sp_material);
    out[2] = (int32_t)sizeof(sp_primitive); out[3] = (int32_t)sizeof(sp_collider);
    out[4] = (int32_t)sizeof(sp_light);    out[5] = (int32_t)sizeof(sp_stats);
    return 0;
}

const char* sp_last_error(void) { return g_error.c_str(); }

int sp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static int init_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail("no CUDA device available (%s); sightpy-b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n || device >= SP_MAX_DEVICES) return fail("device %d out of range (0..%d)", device, std::min(n, SP_MAX_DEVICES) - 1);
    SP_ENTER(device);
    if (g_ctx[device].ready) return 0;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail("device %d is %s (sm_%d%d); this library is built for sm_100a (B200) only", device, prop.name,
                    prop.major, prop.minor);
    float plain[256], linear[256];
    for (int b = 0; b < 256; ++b) {
        const double x = b / 256.0;                                   // image_functions.py:7-9
        plain[b] = (float)x;                                          // colour_functions.py:21-28
        linear[b] = (float)(x <= 0.03928 ? x / 12.92 : std::pow((x + 0.055) / 1.055, 2.4));
    }
    CUDA_TRY(sp_upload_decode_tables(plain, linear));
    g_ctx[device].ready = true;
    return 0;
}

int sp_init(int device) {
    if (int rc = init_device(device)) return rc;
    g_device = device;
    cudaSetDevice(device);
    return 0;
}

// Bind the process to several GPUs of the node: device_ids[0] becomes the default device (new scenes, the one that
// gathers and resolves group frames) and is given peer access to the others, so that it can read their frames over
// NVLink (sp_render_group).
int sp_init_devices(int n, const int* device_ids) {
    if (n < 1 || !device_ids) return fail("sp_init_devices: need at least one device id");
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j)
            if (device_ids[i] == device_ids[j]) return fail("sp_init_devices: device %d listed twice", device_ids[i]);
    for (int i = 0; i < n; ++i)
        if (int rc = init_device(device_ids[i])) return rc;
    g_device = device_ids[0];
    SP_ENTER(g_device);
    for (int i = 1; i < n; ++i) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, g_device, device_ids[i]) == cudaSuccess && can) {
            cudaError_t e = cudaDeviceEnablePeerAccess(device_ids[i], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail("cudaDeviceEnablePeerAccess(%d): %s", device_ids[i], cudaGetErrorString(e));
            cudaGetLastError();
        }
    }
    return 0;
}

int sp_default_device(void) { return g_device; }

static void release_device_ctx(int device) {
    SP_ENTER(device);
    DeviceCtx& c = g_ctx[device];
    if (!c.ready) return;
    cudaDeviceSynchronize();
    for (auto& t : c.tex_cache) cudaFree(t.d);
    c.tex_cache.clear();
    pool_flush();
    for (auto st : c.stream_pool) cudaStreamDestroy(st);
    c.stream_pool.clear();
    for (auto e : c.event_pool) cudaEventDestroy(e);
    c.event_pool.clear();
    c.ready = false;
}

void sp_shutdown(void) {
    for (int d = 0; d < SP_MAX_DEVICES; ++d) release_device_ctx(d);
    g_device = -1;
}

// Return idle pooled device memory (wavefront queues of destroyed / re-committed scenes, unreferenced cached
// textures) to the driver, on every bound device.
void sp_trim(void) {
    for (int d = 0; d < SP_MAX_DEVICES; ++d) {
        if (!g_ctx[d].ready) continue;
        SP_ENTER(d);
        pool_flush();
        auto& tc = g_ctx[d].tex_cache;
        for (size_t i = tc.size(); i-- > 0;)
            if (tc[i].refs == 0) { cudaFree(tc[i].d); tc.erase(tc.begin() + (long)i); }
    }
}

int sp_scene_create_on(sp_scene** out, int device) {
    if (!out) return fail("sp_scene_create: null output pointer");
    if (device < 0 || device >= SP_MAX_DEVICES || !g_ctx[device].ready) return fail("sp_scene_create: device %d is not initialised (sp_init / sp_init_devices)", device);
    *out = new sp_scene();
    (*out)->device = device;
    return 0;
}

int sp_scene_create(sp_scene** out) {
    if (g_device < 0) return fail("sp_scene_create: call sp_init first");
    return sp_scene_create_on(out, g_device);
}

void sp_scene_destroy(sp_scene* s) {
    if (!s) return;
    SP_ENTER(s->device);
    delete s;
}

// =================================================================================================
// scene description
// =================================================================================================
#define NEED_SCENE(s) do { if (!(s)) return fail("%s: null scene", __func__); (s)->committed = false; } while (0)

int sp_scene_set_globals(sp_scene* s, const double ambient[3], const double* media_re, const double* media_im,
                         int n_media) {
    NEED_SCENE(s);
    if (n_media < 1 || n_media > 255) return fail("sp_scene_set_globals: need 1..255 media (row 0 = scene.n)");
    memcpy(s->ambient, ambient, sizeof s->ambient);
    s->media_re.assign(media_re, media_re + 3 * (size_t)n_media);
    s->media_im.assign(media_im, media_im + 3 * (size_t)n_media);
    return 0;
}

int sp_scene_set_camera(sp_scene* s, const sp_camera* cam) {
    NEED_SCENE(s);
    if (!cam || cam->width < 1 || cam->height < 1) return fail("sp_scene_set_camera: invalid camera");
    if ((uint64_t)cam->width * (uint64_t)cam->height > 0x7FFFFFFFull) return fail("sp_scene_set_camera: frame too large");
    s->cam = *cam;
    s->has_camera = true;
    return 0;
}

// what primary-ray generation would otherwise divide by, per ray (sp_camera_ray)
static void derive_camera(DCamera& c) {
    c.w_magic = c.W > 1 ? (~0ull) / (unsigned long long)c.W + 1ull : 0ull;     // W == 1 is special-cased on the device
    c.step_x = c.W > 1 ? (float)((double)c.cam_w / (double)(c.W - 1)) : 0.f;
    c.step_y = c.H > 1 ? (float)((double)c.cam_h / (double)(c.H - 1)) : 0.f;
    c.jitter_x = (float)((double)c.cam_w / (double)c.W);
    c.jitter_y = (float)((double)c.cam_h / (double)c.H);
}

int sp_scene_update_camera(sp_scene* s, const sp_camera* cam) {
    SP_ENTER(s ? s->device : (g_device >= 0 ? g_device : 0));
    if (!s || !cam) return fail("sp_scene_update_camera: invalid arguments");
    if (!s->committed || !s->has_camera) return fail("sp_scene_update_camera: scene not committed with a camera");
    if (cam->width != s->cam.width || cam->height != s->cam.height)
        return fail("sp_scene_update_camera: the frame size changed (%dx%d -> %dx%d): set the camera and commit again",
                    s->cam.width, s->cam.height, cam->width, cam->height);
    s->cam = *cam;
    DCamera& c = s->d.cam;
    c.look_from = f3(cam->look_from); c.right = f3(cam->right); c.up = f3(cam->up); c.fwd = f3(cam->fwd);
    c.cam_w = (float)cam->cam_w; c.cam_h = (float)cam->cam_h; c.lens_radius = (float)cam->lens_radius;
    c.focal_distance = (float)cam->focal_distance;
    derive_camera(c);
    return 0;
}

int sp_scene_clear_textures(sp_scene* s) {
    SP_ENTER(s ? s->device : (g_device >= 0 ? g_device : 0));
    NEED_SCENE(s);
    s->textures.clear();
    return 0;
}

int sp_scene_add_texture_blurred(sp_scene* s, uint64_t key, const uint8_t* rgb, int H, int W, int decode, double cube_blur,
                                 int* tex_id) {
    SP_ENTER(s ? s->device : (g_device >= 0 ? g_device : 0));
    NEED_SCENE(s);
    if (!rgb || H < 1 || W < 1) return fail("sp_scene_add_texture: invalid image");
    if (decode != SP_DECODE_PLAIN && decode != SP_DECODE_LINEAR) return fail("sp_scene_add_texture: unknown decode %d", decode);
    if (!(cube_blur >= 0.0)) return fail("sp_scene_add_texture_blurred: blur must be >= 0");
    if (cube_blur > 0.0 && (H / 3 < 1 || 4 * (H / 3) > W))
        return fail("sp_scene_add_texture_blurred: a %dx%d image is not a 3 x 4 cross of square faces", W, H);
    HostTexture t;
    t.H = H; t.W = W; t.decode = decode; t.key = key; t.cube_blur = cube_blur;
    const int cached = key != 0 ? tex_cache_find(key, H, W, decode) : -1;
    if (cached >= 0 && !s->holds_texture(key)) {                       // resident: pin it for the life of this scene
        ctx().tex_cache[(size_t)cached].refs++;
        s->cached_tex_keys.push_back(key);
    }
    if (cached < 0) {                                                  // not resident yet: pack RGB8 -> one word per texel
        t.texels.resize((size_t)H * W);
        for (size_t i = 0; i < t.texels.size(); ++i)
            t.texels[i] = (uint32_t)rgb[3 * i] | ((uint32_t)rgb[3 * i + 1] << 8) | ((uint32_t)rgb[3 * i + 2] << 16);
    }
    s->textures.push_back(std::move(t));
    if (tex_id) *tex_id = (int)s->textures.size() - 1;
    return 0;
}

int sp_scene_add_texture_keyed(sp_scene* s, uint64_t key, const uint8_t* rgb, int H, int W, int decode, int* tex_id) {
    return sp_scene_add_texture_blurred(s, key, rgb, H, W, decode, 0.0, tex_id);
}

int sp_scene_add_texture(sp_scene* s, const uint8_t* rgb, int H, int W, int decode, int* tex_id) {
    return sp_scene_add_texture_blurred(s, 0, rgb, H, W, decode, 0.0, tex_id);
}

int sp_scene_set_materials(sp_scene* s, const sp_material* m, int n) {
    NEED_SCENE(s);
    if (n < 0 || (n > 0 && !m)) return fail("sp_scene_set_materials: invalid arguments");
    s->mats.assign(m, m + n);
    return 0;
}

int sp_scene_set_primitives(sp_scene* s, const sp_primitive* p, int n) {
    NEED_SCENE(s);
    if (n < 0 || (n > 0 && !p)) return fail("sp_scene_set_primitives: invalid arguments");
    s->prims.assign(p, p + n);
    return 0;
}

int sp_scene_set_colliders(sp_scene* s, const sp_collider* c, int n) {
    NEED_SCENE(s);
    if (n < 0 || (n > 0 && !c)) return fail("sp_scene_set_colliders: invalid arguments");
    if (n > 16382) return fail("sp_scene_set_colliders: at most 16382 colliders (ray records carry a 14-bit source id)");
    s->cols.assign(c, c + n);
    return 0;
}

int sp_scene_set_lights(sp_scene* s, const sp_light* l, int n) {
    NEED_SCENE(s);
    if (n < 0 || (n > 0 && !l)) return fail("sp_scene_set_lights: invalid arguments");
    if (n > SP_MAX_LIGHTS) return fail("sp_scene_set_lights: at most %d lights", SP_MAX_LIGHTS);
    s->lights.assign(l, l + n);
    return 0;
}

int sp_scene_set_importance(sp_scene* s, const int32_t* ids, int n) {
    NEED_SCENE(s);
    if (n < 0 || (n > 0 && !ids)) return fail("sp_scene_set_importance: invalid arguments");
    if (n > SP_MAX_IMPORTANCE) return fail("sp_scene_set_importance: at most %d importance-sampled primitives", SP_MAX_IMPORTANCE);
    s->importance.assign(ids, ids + n);
    return 0;
}

int sp_scene_set_shadow_colliders(sp_scene* s, const int32_t* ids, int n) {
    NEED_SCENE(s);
    if (n < 0 || (n > 0 && !ids)) return fail("sp_scene_set_shadow_colliders: invalid arguments");
    s->shadow_ids.assign(ids, ids + n);
    return 0;
}

int sp_scene_commit(sp_scene* s) {
    SP_ENTER(s ? s->device : (g_device >= 0 ? g_device : 0));
    if (!s) return fail("sp_scene_commit: null scene");
    if (!g_ctx[s->device].ready) return fail("sp_scene_commit: call sp_init first");
    if (s->media_re.empty()) return fail("sp_scene_commit: sp_scene_set_globals was not called");
    const int n_tex = (int)s->textures.size(), n_mat = (int)s->mats.size(), n_prim = (int)s->prims.size();
    const int n_col = (int)s->cols.size(), n_media = (int)s->media_re.size() / 3;

    // ---- validation -----------------------------------------------------------------------------
    auto tex_ok = [&](int id) { return id >= -1 && id < n_tex; };
    for (int i = 0; i < n_mat; ++i) {
        const sp_material& m = s->mats[i];
        if (m.kind < SP_MAT_GLOSSY || m.kind > SP_MAT_SKYBOX) return fail("material %d: unknown kind %d", i, m.kind);
        if (!tex_ok(m.normalmap_tex) || !tex_ok(m.color_tex) || !tex_ok(m.aux_tex0) || !tex_ok(m.aux_tex1))
            return fail("material %d: texture id out of range", i);
        if (m.kind == SP_MAT_REFRACTIVE && (m.medium < 0 || m.medium >= n_media)) return fail("material %d: medium out of range", i);
        if (m.kind == SP_MAT_THINFILM && (m.aux_tex0 < 0 || (m.noise_factor != 0.0 && m.aux_tex1 < 0)))
            return fail("material %d: thin film needs its reflectance LUT (and noise) texture", i);
        if (m.kind == SP_MAT_SKYBOX && (m.color_tex < 0 || (m.light_intensity != 0.0 && m.aux_tex0 < 0)))
            return fail("material %d: sky box needs its environment (and light map) texture", i);
        if (m.kind == SP_MAT_DIFFUSE && (m.diffuse_rays < 1 || m.diffuse_rays > 4096 || m.max_diffuse_reflections < 0 ||
                                         m.max_diffuse_reflections > 3))
            return fail("material %d: diffuse_rays must be 1..4096 and max_diffuse_reflections 0..3", i);
    }
    for (int i = 0; i < n_prim; ++i) {
        if (s->prims[i].material < 0 || s->prims[i].material >= n_mat) return fail("primitive %d: material out of range", i);
        if (s->prims[i].max_ray_depth < 0 || s->prims[i].max_ray_depth > 40) return fail("primitive %d: max_ray_depth must be 0..40", i);
    }
    for (int i = 0; i < n_col; ++i) {
        const sp_collider& c = s->cols[i];
        if (c.type < SP_COLLIDER_SPHERE || c.type > SP_COLLIDER_TRIANGLE) return fail("collider %d: unknown type %d", i, c.type);
        if (c.primitive < 0 || c.primitive >= n_prim) return fail("collider %d: primitive out of range", i);
        const sp_material& m = s->mats[s->prims[c.primitive].material];
        const bool uses_uv = (m.color_tex >= 0 && m.kind != SP_MAT_SKYBOX) || m.normalmap_tex >= 0 ||
                             m.kind == SP_MAT_THINFILM || m.kind == SP_MAT_SKYBOX;
        if (c.type == SP_COLLIDER_TRIANGLE && uses_uv) return fail("collider %d: triangles have no uv mapping (triangle.py:79-83)", i);
        if (m.normalmap_tex >= 0 && c.type != SP_COLLIDER_PLANE && c.type != SP_COLLIDER_CUBOID)
            return fail("collider %d: normal maps need a plane or cuboid tangent frame (material.py:32)", i);
    }
    for (int id : s->importance) if (id < 0 || id >= n_prim) return fail("importance list: primitive %d out of range", id);
    for (int id : s->shadow_ids) if (id < 0 || id >= n_col) return fail("shadow list: collider %d out of range", id);

    s->release_device();
    if (!ctx().stream_pool.empty()) { s->own_stream = ctx().stream_pool.back(); ctx().stream_pool.pop_back(); }
    else CUDA_TRY(cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking));
    s->stream = s->user_stream_set ? s->user_stream : s->own_stream;
    DScene& d = s->d;
    memset(&d, 0, sizeof d);

    // ---- materials, fan classes --------------------------------------------------------------------
    d.n_fan_classes = 1;
    d.fan_mult[0] = 1;
    std::vector<DMaterial> dm((size_t)n_mat);
    int max_dr = 0;
    for (int i = 0; i < n_mat; ++i) {
        const sp_material& m = s->mats[i];
        DMaterial& o = dm[i];
        memset(&o, 0, sizeof o);
        o.kind = m.kind; o.medium = m.medium; o.normalmap_tex = m.normalmap_tex; o.color_tex = m.color_tex;
        o.aux_tex0 = m.aux_tex0; o.aux_tex1 = m.aux_tex1; o.diffuse_rays = m.diffuse_rays;
        o.max_dr = m.max_diffuse_reflections; o.index_h = m.index_h; o.index_w = m.index_w;
        o.normalmap_repeat = (float)m.normalmap_repeat; o.color_repeat = (float)m.color_repeat;
        o.color = f3(m.color); o.n_re = f3(m.n_re); o.n_im = f3(m.n_im);
        o.roughness = (float)m.roughness; o.spec_coeff = (float)m.spec_coeff; o.diff_coeff = (float)m.diff_coeff;
        o.thickness = (float)m.thickness; o.noise_factor = (float)m.noise_factor;
        o.ambient_weight = (float)m.ambient_weight; o.light_intensity = (float)m.light_intensity;
        // texel-addressed materials evaluate the hit in double (sp_surface.cuh)
        o.precise = (m.color_tex >= 0 || m.normalmap_tex >= 0 || m.kind == SP_MAT_THINFILM || m.kind == SP_MAT_SKYBOX) ? 1 : 0;
        o.fan_class = 0;
        if (m.kind == SP_MAT_DIFFUSE) {
            max_dr = std::max(max_dr, m.max_diffuse_reflections);
            if (m.diffuse_rays > 1) {
                int cls = -1;
                for (int c = 1; c < d.n_fan_classes; ++c) if (d.fan_mult[c] == m.diffuse_rays) cls = c;
                if (cls < 0) {
                    if (d.n_fan_classes == SP_MAX_FAN_CLASSES)
                        return fail("at most %d distinct diffuse_rays values (> 1) per scene", SP_MAX_FAN_CLASSES - 1);
                    cls = d.n_fan_classes++;
                    d.fan_mult[cls] = m.diffuse_rays;
                }
                o.fan_class = cls;
            }
        }
    }
    for (int c = 0; c < SP_MAX_FAN_CLASSES; ++c) {
        const unsigned long long mult = (unsigned long long)std::max(d.fan_mult[c], 1);
        d.fan_magic[c] = mult == 1 ? 0ull : (~0ull) / mult + 1ull;         // ceil(2^64 / mult); mult == 1 is special-cased below
    }
    CUDA_TRY(s->d_mats.upload(dm));

    std::vector<DPrimitive> dp((size_t)n_prim);
    int max_depth = 0;
    for (int i = 0; i < n_prim; ++i) {
        dp[i].material = s->prims[i].material; dp[i].max_ray_depth = s->prims[i].max_ray_depth;
        dp[i].mc = s->prims[i].mc; dp[i].uv_cross = s->prims[i].uv_cross_layout;
        max_depth = std::max(max_depth, s->prims[i].max_ray_depth);
    }
    CUDA_TRY(s->d_prims.upload(dp));
    // deepest ray.depth that can exist: specular children stop at max_ray_depth, every diffuse
    // bounce adds one more level (diffuse.py tests diffuse_reflections, not depth)
    // (a first Diffuse hit always fans out, whatever max_diffuse_reflections says: diffuse.py:34)
    bool any_diffuse = false;
    for (int i = 0; i < n_mat; ++i) any_diffuse = any_diffuse || s->mats[i].kind == SP_MAT_DIFFUSE;
    s->n_levels = std::min(max_depth + (any_diffuse ? std::max(max_dr, 1) : 0) + 1, SP_MAX_LEVELS - 1);

    std::vector<DCollider> dc((size_t)n_col);
    const int PL = 44;                                   // SP_DEV_PAYLOAD: 40 ABI slots + derived reciprocals
    std::vector<double> dcd((size_t)n_col * PL, 0.0);
    for (int i = 0; i < n_col; ++i) {
        dc[i].type = s->cols[i].type; dc[i].prim = s->cols[i].primitive;
        double* pd = &dcd[(size_t)i * PL];
        for (int k = 0; k < 40; ++k) pd[k] = s->cols[i].p[k];
        switch (s->cols[i].type) {
        case SP_COLLIDER_SPHERE: pd[4] = 1.0 / pd[3]; break;                                   // 1 / radius
        case SP_COLLIDER_PLANE: pd[25] = 1.0 / pd[12]; pd[26] = 1.0 / pd[13]; break;           // 1 / w, 1 / h
        case SP_COLLIDER_CUBOID: for (int k = 0; k < 3; ++k) pd[40 + k] = 1.0 / pd[18 + k]; break;   // 1 / (width, height, length)
        default: break;
        }
        for (int k = 0; k < PL; ++k) dc[i].p[k] = (float)pd[k];
    }
    CUDA_TRY(s->d_cols.upload(dc));
    std::vector<DColInfo> dinfo((size_t)n_col);
    for (int i = 0; i < n_col; ++i) {
        const sp_primitive& pr = s->prims[s->cols[i].primitive];
        const sp_material& m = s->mats[pr.material];
        DColInfo& ci = dinfo[i];
        ci.type = (uint8_t)s->cols[i].type; ci.kind = (uint8_t)m.kind; ci.mc = pr.mc ? 1 : 0;
        ci.fan_class = (uint8_t)dm[pr.material].fan_class;
        ci.max_ray_depth = (int16_t)pr.max_ray_depth; ci.max_dr = (int16_t)m.max_diffuse_reflections;
        ci.slot = 0xFFFFFFFFu; ci.w_cos = (float)m.ambient_weight;
    }
    std::vector<float4> dlite((size_t)n_col);
    for (int i = 0; i < n_col; ++i) {
        const DMaterial& m = dm[s->prims[s->cols[i].primitive].material];
        dlite[i] = make_float4(m.color.x, m.color.y, m.color.z, m.diffuse_rays > 0 ? 1.f / (float)m.diffuse_rays : 1.f);
    }
    CUDA_TRY(s->d_collite.upload(dlite));
    CUDA_TRY(s->d_cols_d.upload(dcd));

    // ---- textures ---------------------------------------------------------------------------------------
    std::vector<DTexture> td((size_t)n_tex);
    s->d_texels.resize((size_t)n_tex);
    for (int i = 0; i < n_tex; ++i) {
        HostTexture& ht = s->textures[i];
        if (ht.key != 0) {
            int ci = tex_cache_find(ht.key, ht.H, ht.W, ht.decode);
            if (ci < 0) {
                if (ht.texels.empty()) return fail("texture %d: key %llu left the cache between description and commit", i, (unsigned long long)ht.key);
                CachedTexture ct{ht.key, ht.H, ht.W, ht.decode, nullptr, ht.texels.size() * sizeof(uint32_t), 0, 0};
                cudaError_t em = cudaMalloc(&ct.d, ct.bytes);
                if (em != cudaSuccess && !ctx().pool.empty()) {          // make room (idle pooled buffers) and retry once
                    cudaGetLastError();
                    pool_flush();
                    em = cudaMalloc(&ct.d, ct.bytes);
                }
                CUDA_TRY(em);
                CUDA_TRY(cudaMemcpy(ct.d, ht.texels.data(), ct.bytes, cudaMemcpyHostToDevice));
                if (ht.cube_blur > 0.0)
                    if (int rcb = blur_texture_in_place(ct.d, ht.H, ht.W, ht.cube_blur)) { cudaFree(ct.d); return rcb; }
                ctx().tex_cache.push_back(ct);
                ci = (int)ctx().tex_cache.size() - 1;
            }
            if (!s->holds_texture(ht.key)) { ctx().tex_cache[(size_t)ci].refs++; s->cached_tex_keys.push_back(ht.key); }
            ctx().tex_cache[(size_t)ci].last_use = ++ctx().tex_clock;
            td[i].texels = ctx().tex_cache[(size_t)ci].d;
        } else {
            CUDA_TRY(s->d_texels[i].upload(ht.texels));
            if (ht.cube_blur > 0.0)
                if (int rcb = blur_texture_in_place(s->d_texels[i].p, ht.H, ht.W, ht.cube_blur)) return rcb;
            td[i].texels = s->d_texels[i].p;
        }
        td[i].H = ht.H; td[i].W = ht.W;
        td[i].decode = s->textures[i].decode; td[i].pad = 0;
    }
    CUDA_TRY(s->d_texdesc.upload(td));

    // ---- media (complex index + Beer-Lambert coefficient, refractive.py:113-121) -------------------------
    std::vector<DMedium> med((size_t)n_media);
    const double lambda[3] = {630.0, 550.0, 475.0};
    for (int i = 0; i < n_media; ++i) {
        med[i].re = f3(&s->media_re[3 * (size_t)i]);
        med[i].im = f3(&s->media_im[3 * (size_t)i]);
        double ab[3];
        for (int c = 0; c < 3; ++c) ab[c] = 2.0 * s->media_im[3 * (size_t)i + c] * 2.0 * M_PI / lambda[c] * 1e9;
        med[i].absorb = f3(ab);
        const double* re = &s->media_re[3 * (size_t)i];
        const double* im = &s->media_im[3 * (size_t)i];
        med[i].grey = ((float)re[0] == (float)re[1] && (float)re[1] == (float)re[2] && re[0] > 0.0 &&
                       std::fabs(im[0]) <= 1e-4 * re[0] && std::fabs(im[1]) <= 1e-4 * re[0] && std::fabs(im[2]) <= 1e-4 * re[0]) ? 1 : 0;
    }
    CUDA_TRY(s->d_media.upload(med));

    // ---- geometry streams -------------------------------------------------------------------------------
    // large scenes: the small colliders go into a BVH, the streams keep the scene-sized ones
    std::vector<Aabb> boxes((size_t)n_col);
    for (int i = 0; i < n_col; ++i) boxes[i] = collider_aabb(s->cols[i]);
    std::vector<char> in_bvh = bvh_membership(s->cols, boxes);
    if (!s->opt_bvh) std::fill(in_bvh.begin(), in_bvh.end(), 0);     // option "bvh" = 0: exhaustive loop over staged chunks
    std::vector<char> casts_shadow((size_t)n_col, 0);
    for (int id : s->shadow_ids) casts_shadow[id] = 1;
    const BuiltBvh bvh = build_bvh(s->cols, in_bvh, boxes, casts_shadow);
    CUDA_TRY(s->bvh_nodes.upload(bvh.nodes));
    CUDA_TRY(s->bvh_items.upload(bvh.items));
    CUDA_TRY(s->bvh_data.upload(bvh.data));
    d.bvh.nodes = s->bvh_nodes.p; d.bvh.items = s->bvh_items.p; d.bvh.data = s->bvh_data.p;
    d.bvh.n_nodes = (int)bvh.nodes.size() / 4; d.bvh.n_items = (int)bvh.items.size();
    d.n_shadow_casters = (int)s->shadow_ids.size();
    std::vector<int32_t> all_ids, shadow_stream_ids;
    for (int i = 0; i < n_col; ++i) if (!in_bvh[i]) all_ids.push_back(i);
    for (int id : s->shadow_ids) if (!in_bvh[id]) shadow_stream_ids.push_back(id);
    BuiltStream all = build_stream(s->cols, all_ids), shadow = build_stream(s->cols, shadow_stream_ids);
    for (int i = 0; i < n_col; ++i) {                    // (chunk, type << 28 | local) -> chunk << 24 | type << 20 | local
        const int2 w = all.slot[i];
        dinfo[i].slot = w.x < 0 ? 0xFFFFFFFFu        // inside the BVH: recognised there by its collider id
                                : ((uint32_t)w.x << 24) | (((uint32_t)w.y >> 28) << 20) | ((uint32_t)w.y & 0xFFFFFu);
    }
    if (all.chunk_off.size() - 1 > 255) return fail("too many geometry chunks");
    CUDA_TRY(s->d_colinfo.upload(dinfo));
    CUDA_TRY(s->geom_all.upload(all.data));
    CUDA_TRY(s->off_all.upload(all.chunk_off));
    CUDA_TRY(s->geom_shadow.upload(shadow.data));
    CUDA_TRY(s->off_shadow.upload(shadow.chunk_off));
    CUDA_TRY(s->slot_shadow.upload(shadow.slot));
    d.all.data = s->geom_all.p; d.all.chunk_off = s->off_all.p;
    d.all.n_chunks = (int)all.chunk_off.size() - 1; d.all.n_items = all.n_items;
    for (size_t c = 0; c + 1 < all.chunk_off.size(); ++c)
        d.all.max_chunk_vec4 = std::max(d.all.max_chunk_vec4, all.chunk_off[c + 1] - all.chunk_off[c]);
    d.shadow.data = s->geom_shadow.p; d.shadow.chunk_off = s->off_shadow.p;
    d.shadow.n_chunks = (int)shadow.chunk_off.size() - 1; d.shadow.n_items = shadow.n_items;

    d.colliders = s->d_cols.p; d.col_info = s->d_colinfo.p; d.col_lite = s->d_collite.p; d.colliders_d = s->d_cols_d.p; d.prims = s->d_prims.p; d.mats = s->d_mats.p;
    d.textures = s->d_texdesc.p; d.media = s->d_media.p;
    d.n_lights = (int)s->lights.size();
    for (int i = 0; i < d.n_lights; ++i) {
        d.lights[i].kind = s->lights[i].kind; d.lights[i].vec = f3(s->lights[i].vec); d.lights[i].color = f3(s->lights[i].color);
    }
    d.n_importance = (int)s->importance.size();
    d.inv_n_importance = d.n_importance ? 1.f / (float)d.n_importance : 0.f;
    for (int i = 0; i < d.n_importance; ++i) {
        const sp_primitive& p = s->prims[s->importance[i]];
        d.importance[i].center = f3(p.center); d.importance[i].radius = (float)p.bounded_sphere_radius;
    }
    d.ambient = f3(s->ambient);
    d.n_colliders = n_col;
    if (s->has_camera) {
        const sp_camera& c = s->cam;
        d.cam.look_from = f3(c.look_from); d.cam.right = f3(c.right); d.cam.up = f3(c.up); d.cam.fwd = f3(c.fwd);
        d.cam.cam_w = (float)c.cam_w; d.cam.cam_h = (float)c.cam_h; d.cam.lens_radius = (float)c.lens_radius;
        d.cam.focal_distance = (float)c.focal_distance; d.cam.W = c.width; d.cam.H = c.height;
        derive_camera(d.cam);
        CUDA_TRY(s->accum.alloc((size_t)c.width * c.height));
        CUDA_TRY(cudaMemset(s->accum.p, 0, s->accum.n * sizeof(float4)));
    }
    for (auto& sl : s->slot) {
        CUDA_TRY(sl.counts.alloc((size_t)(SP_MAX_LEVELS + 1) * SP_COUNTS_PER_LEVEL));
        CUDA_TRY(sl.d_stats.alloc(1));
        if (!sl.h_counts) CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&sl.h_counts), (size_t)(SP_MAX_LEVELS + 1) * SP_COUNTS_PER_LEVEL * sizeof(uint32_t), cudaHostAllocDefault));
        if (!sl.h_stats) CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&sl.h_stats), sizeof(DeviceStats), cudaHostAllocDefault));
        sl.events.resize((size_t)s->n_levels + 1);
        auto take_event = [&](cudaEvent_t& e) -> cudaError_t {
            if (!ctx().event_pool.empty()) { e = ctx().event_pool.back(); ctx().event_pool.pop_back(); return cudaSuccess; }
            return cudaEventCreate(&e);
        };
        for (auto& e : sl.events) CUDA_TRY(take_event(e));
        if (!sl.done) CUDA_TRY(take_event(sl.done));
        sl.busy = false;
    }
    uint32_t needed = 0;
    for (int i = 0; i < n_col; ++i) {
        const sp_material& m = s->mats[s->prims[s->cols[i].primitive].material];
        if (m.color_tex >= 0 || m.normalmap_tex >= 0) needed |= SP_F_TEX;
        switch (m.kind) {
        case SP_MAT_GLOSSY: needed |= SP_F_GLOSSY; break;
        case SP_MAT_REFRACTIVE: needed |= SP_F_REFR; break;
        case SP_MAT_THINFILM: needed |= SP_F_THIN | SP_F_TEX; break;
        case SP_MAT_DIFFUSE: needed |= SP_F_DIFFUSE; break;
        case SP_MAT_SKYBOX: needed |= SP_F_SKY | SP_F_TEX; break;
        default: break;
        }
    }
    if (d.bvh.n_nodes > 0) needed |= SP_F_BVH;
    s->material_set = sp_pick_material_set(needed);
    s->d.use_warp_kernel = s->opt_warp ? 1 : 0;
    s->d.use_split = s->opt_split ? 1 : 0;
    s->kind_mask = 0;
    for (const auto& m : s->mats) s->kind_mask |= 1u << (m.kind & 31);
    if (int rc = pick_kernels(s)) return rc;
    const uint64_t sig = shape_signature(s);
    if (sig != s->shape_sig) {                   // a different scene: measure the queue occupancy afresh
        s->use_ray = s->use_fan = 0.0;
        s->chunk_limit = 0;
        s->shape_sig = sig;
    }
    s->committed = true;
    return 0;
}

// =================================================================================================
// wavefront driver
// =================================================================================================
// Primaries per chunk and records per queue when the caller sets neither: queues hold 24 records per primary of a
// chunk (a diffuse first bounce turns 1 primary into diffuse_rays secondary hits): 192 Mi records each (54 GB with two
// fan classes) for a full 8 Mi-primary chunk, proportionally less for small jobs; options "chunk_primaries",
// "ray_queue_capacity" and "fan_queue_capacity" trade memory for launch size (4 Mi-primary chunks: 27 GB, -1 % on the
// headline frame).  The driver measures the real occupancy on a first chunk and sizes later chunks to fit; a chunk
// that overflows a queue all the same is rendered again at half the size (see the render loop below).
#define SP_DEFAULT_CHUNK ((int64_t)8 << 20)      /* measured 4 Mi -> 35.3, 8 Mi -> 35.7, 16 Mi -> 35.8 Grays/s */

static int64_t default_chunk(const sp_scene* s) { return s->opt_chunk > 0 ? s->opt_chunk : SP_DEFAULT_CHUNK; }

static int ensure_queues(sp_scene* s, uint64_t primaries) {
    const int64_t chunk = default_chunk(s);
    int64_t auto_cap = 24 * (int64_t)std::min<uint64_t>(std::max<uint64_t>(primaries, 1), (uint64_t)chunk);
    auto_cap = std::max<int64_t>((auto_cap + 0xFFFFF) & ~(int64_t)0xFFFFF, (int64_t)1 << 20);
    uint32_t want_ray = (uint32_t)std::min<int64_t>(s->opt_ray_cap > 0 ? s->opt_ray_cap : auto_cap, 0x7FFFFFF0ll);
    uint32_t want_fan = (uint32_t)std::min<int64_t>(s->opt_fan_cap > 0 ? s->opt_fan_cap : auto_cap, 0x7FFFFFF0ll);
    for (int c = 0; c < s->d.n_fan_classes; ++c)       // the kernels count work items in 32 bits
        want_fan = (uint32_t)std::min<uint64_t>(want_fan, 0x7FFFFFFFull / (uint64_t)std::max(s->d.fan_mult[c], 1));
    if (s->opt_ray_cap == 0 && s->ray_cap >= want_ray) want_ray = s->ray_cap;       // never shrink on our own
    if (s->opt_fan_cap == 0 && s->fan_cap >= want_fan) want_fan = s->fan_cap;
    if (want_ray != s->ray_cap) {
        for (int i = 0; i < 2; ++i) CUDA_TRY(s->ray_q[i].alloc(want_ray));
        s->ray_cap = want_ray;
    }
    if (want_fan != s->fan_cap || s->fan_q[0].n != (size_t)want_fan * s->d.n_fan_classes) {
        for (int i = 0; i < 2; ++i) CUDA_TRY(s->fan_q[i].alloc((size_t)want_fan * s->d.n_fan_classes));
        s->fan_cap = want_fan;
    }
    return 0;
}

// Enqueue all levels of one chunk into slot k: the level launches, then (render calls) the fold of the chunk's scratch
// frame into the accumulation buffer, the copies of its counters into pinned host memory, and the slot's event.
static int enqueue_chunk(sp_scene* s, const ChunkJob& job, int k, float4* fold_into, float4* fold_from, uint32_t fold_n) {
    ChunkSlot& sl = s->slot[k];
    int n_levels = (job.run == SP_RUN_FULL) ? s->n_levels : 1;
    if (s->opt_max_levels > 0) n_levels = std::min<int>(n_levels, (int)s->opt_max_levels);   // debugging aid
    const int ncl = SP_COUNTS_PER_LEVEL;
    const bool warp = job.run == SP_RUN_FULL && sp_use_warp_kernel(s->d, s->material_set);
    // scenes behind a BVH: the nearest hits of every level are found by a kernel of their own ahead of the level launch
    const bool pretrace = job.run == SP_RUN_FULL && s->opt_pretrace && sp_can_pretrace(s->d, s->material_set);
    if (pretrace) {
        // items of the widest level: primaries, or queued records times their fan size (from the occupancy seen so far)
        double per_primary = 1.0 + s->use_ray;
        for (int c = 0; c < s->d.n_fan_classes; ++c) per_primary = std::max(per_primary, s->use_ray + s->use_fan * s->d.fan_mult[c] * s->d.n_fan_classes);
        if (s->use_ray == 0.0 && s->use_fan == 0.0) per_primary = 64.0;
        const size_t want = (size_t)std::min<double>(std::max<double>(1.3 * per_primary * job.n_items, 1 << 20), (double)((size_t)1 << 30));
        if (s->d_hits.n < want) CUDA_TRY(s->d_hits.alloc(want));
        // shadow-ray requests of one level: at most one per item and light; a quarter of the items' worth is kept (a
        // request that finds the queue full is traversed inside the level kernel instead)
        if (s->d.n_lights > 0 && s->d.n_shadow_casters > 0) {
            const size_t want_sh = std::min<size_t>(std::max<size_t>(want / 4, (size_t)1 << 20), (size_t)128 << 20);
            if (s->d_shq.n < 3 * want_sh) CUDA_TRY(s->d_shq.alloc(3 * want_sh));
            if (s->d_shq_count.n < 2 * (size_t)SP_MAX_LEVELS) CUDA_TRY(s->d_shq_count.alloc(2 * (size_t)SP_MAX_LEVELS));
            CUDA_TRY(cudaMemsetAsync(s->d_shq_count.p, 0, s->d_shq_count.n * sizeof(uint32_t), s->stream));
        }
    }
    const bool defer_shadows = pretrace && s->d_shq.p && s->d.n_lights > 0 && s->d.n_shadow_casters > 0;
    // Whitted scenes: hit kernel + per-material shade kernels (option "split_kernels": 1 = level 0 of launches of at
    // least 256 Ki primaries, 2 = every level); per-item hit records and per-kind item lists sized for the widest level
    // that runs this way (the primaries, or every queued record)
    static const int env_split = [] { const char* e = getenv("SIGHTPY_SPLIT"); return (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : -1; }();
    const int split_mode = (job.run == SP_RUN_FULL && sp_can_split(s->d, s->material_set)) ? (env_split >= 0 ? env_split : (int)s->opt_split) : 0;
    const bool split = split_mode == 2 || (split_mode == 1 && job.n_items >= (1u << 18));
    size_t split_cap = 0;
    if (split) {
        // level 0 only: the largest chunk this call can cut (so that the buffers are allocated by the first chunk of the
        // first frame, not again whenever a later chunk is larger)
        const size_t largest = (size_t)std::min<uint64_t>(std::max<uint64_t>(s->call_primaries, job.n_items), 8ull * (uint64_t)default_chunk(s));
        split_cap = split_mode == 2 ? std::max<size_t>(largest, s->ray_cap) : largest;
        if (s->d_hits.n < split_cap) CUDA_TRY(s->d_hits.alloc(split_cap));
        if (s->d_klist.n < split_cap * SPS_KINDS_HOST) CUDA_TRY(s->d_klist.alloc(split_cap * SPS_KINDS_HOST));
        if (s->d_kcount.n < 8 * (size_t)SP_MAX_LEVELS) CUDA_TRY(s->d_kcount.alloc(8 * (size_t)SP_MAX_LEVELS));
        CUDA_TRY(cudaMemsetAsync(s->d_kcount.p, 0, s->d_kcount.n * sizeof(uint32_t), s->stream));
    }
    int split_launches = 0;
    CUDA_TRY(cudaMemsetAsync(sl.counts.p, 0, (size_t)(n_levels + 1) * ncl * sizeof(uint32_t), s->stream));
    CUDA_TRY(cudaMemsetAsync(sl.d_stats.p, 0, sizeof(DeviceStats), s->stream));
    for (int L = 0; L < n_levels; ++L) {
        LevelArgs a;
        memset(&a, 0, sizeof a);
        a.level = L; a.run = job.run;
        a.source = (L == 0) ? job.source : SP_SRC_QUEUES;
        a.pix_begin = job.pix_begin; a.n_pix = job.n_pix; a.sample_begin = job.sample_begin;
        a.n_items0 = job.n_items; a.user_base = job.user_base; a.user_o = job.user_o; a.user_d = job.user_d;
        a.n_pix_magic = job.n_pix > 1u ? (~0ull) / (unsigned long long)job.n_pix + 1ull : 0ull;
        a.tiles = job.tiles; a.tile_shift = job.tile_shift; a.tiles_x = job.tiles_x;
        a.in_rays = s->ray_q[L & 1].view(s->ray_cap);
        a.in_fans = s->fan_q[L & 1].view(s->fan_cap * s->d.n_fan_classes);
        a.out.rays = s->ray_q[(L + 1) & 1].view(s->ray_cap);
        a.out.fans = s->fan_q[(L + 1) & 1].view(s->fan_cap * s->d.n_fan_classes);
        for (int c = 0; c < SP_MAX_FAN_CLASSES; ++c) {
            a.in_fan_base[c] = a.out.fan_base[c] = (uint32_t)c * s->fan_cap;
            a.in_fan_cap[c] = a.out.fan_cap[c] = (c < s->d.n_fan_classes) ? s->fan_cap : 0u;
        }
        a.in_counts = sl.counts.p + (size_t)L * ncl;
        a.out.counts = sl.counts.p + (size_t)(L + 1) * ncl;
        a.out.stats = sl.d_stats.p;
        a.accum = job.accum;
        a.out_hit = job.out_hit; a.out_t = job.out_t; a.out_o = job.out_o; a.out_d = job.out_d; a.out_n = job.out_n;
        a.shadow_slot = s->slot_shadow.p;
        a.hits = pretrace ? s->d_hits.p : nullptr;
        a.hits_cap = pretrace ? (uint32_t)std::min<size_t>(s->d_hits.n, 0xFFFFFFFFull) : 0u;
        CUDA_TRY(cudaEventRecord(sl.events[L], s->stream));
        a.shq = defer_shadows ? s->d_shq.p : nullptr;
        a.shq_cap = defer_shadows ? (uint32_t)std::min<size_t>(s->d_shq.n / 3, 0xFFFFFFFFull) : 0u;
        a.shq_count = defer_shadows ? s->d_shq_count.p + 2 * (size_t)L : nullptr;
        if (split && (split_mode == 2 || L == 0)) {
            a.hits = s->d_hits.p; a.hits_cap = (uint32_t)std::min<size_t>(split_cap, 0xFFFFFFFFull);
            a.kind_list = s->d_klist.p; a.kind_cap = a.hits_cap; a.kind_count = s->d_kcount.p + 8 * (size_t)L;
            int n_k = 0;
            CUDA_TRY(sp_launch_split_level(s->d, a, s->kind_mask, s->device, s->stream, &n_k));
            split_launches += n_k;
            continue;
        }
        if (pretrace) CUDA_TRY(sp_launch_trace(s->d, a, s->material_set, s->device, s->stream));
        CUDA_TRY(sp_launch_level(s->d, a, s->material_set, L == 0 ? s->grid0 : s->grid_q, s->stream));
        if (defer_shadows) CUDA_TRY(sp_launch_shadow(s->d, a, s->device, s->stream));
    }
    CUDA_TRY(cudaEventRecord(sl.events[n_levels], s->stream));
    // the chunk's radiance moves from the scratch frame into the accumulation buffer — on the device, unless one of the
    // chunk's queues overflowed (then the scratch frame is cleared instead and the host renders the chunk again)
    sl.folded = fold_into != nullptr;
    if (fold_into) CUDA_TRY(sp_launch_fold(fold_into, fold_from, fold_n, sl.d_stats.p, s->stream));
    CUDA_TRY(cudaMemcpyAsync(sl.h_counts, sl.counts.p, (size_t)(n_levels + 1) * ncl * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaMemcpyAsync(sl.h_stats, sl.d_stats.p, sizeof(DeviceStats), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaEventRecord(sl.done, s->stream));
    sl.busy = true; sl.job = job; sl.n_levels = n_levels; sl.warp = warp; sl.pretrace = pretrace; sl.defer_shadows = defer_shadows;
    sl.split = split; sl.split_mode = split_mode; sl.split_launches = split_launches;
    return 0;
}

// Wait for the chunk in slot k and fold its counters into `st`.  `overflow` comes back true when a queue was too small
// for the chunk: nothing of it reached the frame and the caller renders it again in smaller pieces.
static int finish_chunk(sp_scene* s, int k, sp_stats* st, bool& overflow) {
    ChunkSlot& sl = s->slot[k];
    overflow = false;
    if (!sl.busy) return 0;
    sl.busy = false;
    CUDA_TRY(cudaEventSynchronize(sl.done));
    const ChunkJob& job = sl.job;
    const int n_levels = sl.n_levels, ncl = SP_COUNTS_PER_LEVEL, split_mode = sl.split_mode, split_launches = sl.split_launches;
    const bool warp = sl.warp, pretrace = sl.pretrace, defer_shadows = sl.defer_shadows, split = sl.split;
    const uint32_t* counts = sl.h_counts;
    const DeviceStats& ds = *sl.h_stats;
    uint64_t peak_r = 0, peak_f = 0;
    for (int L = 1; L <= n_levels; ++L) {
        peak_r = std::max<uint64_t>(peak_r, counts[(size_t)L * ncl]);
        for (int c = 0; c < s->d.n_fan_classes; ++c) peak_f = std::max<uint64_t>(peak_f, counts[(size_t)L * ncl + 1 + c]);
    }
    if (ds.overflow >> 16)
        return fail("internal consistency check failed in the wavefront kernels (SP_CHECKED build): code mask 0x%x", ds.overflow >> 16);
    // the kernels raise the flag whenever a reservation does not fit; sp_fold_kernel decides by the same flag
    const bool overflowed = (ds.overflow & 0xFFFFu) != 0u;
    if (!overflowed && (peak_r > s->ray_cap || peak_f > s->fan_cap))
        return fail("internal error: a wavefront queue holds more records than its capacity without the overflow flag");
    if (job.n_items > 0) {
        // a queue that overflowed stopped counting at the level that failed: later levels may have needed more
        const double grow = overflowed ? 1.5 : 1.0;
        s->use_ray = std::max(s->use_ray, grow * (double)peak_r / job.n_items);
        s->use_fan = std::max(s->use_fan, grow * (double)peak_f / job.n_items);
    }
    if (overflowed) {
        overflow = true;
        g_error = "wavefront queue overflow";
        char buf[256];
        snprintf(buf, sizeof buf, "wavefront queue overflow (%llu ray / %llu fan records for %u primaries; capacities %u / %u)",
                 (unsigned long long)peak_r, (unsigned long long)peak_f, job.n_items, s->ray_cap, s->fan_cap);
        g_error = buf;
        return 0;
    }
    if (st) {
        st->chunks += 1;
        if (sl.folded) st->kernel_launches += 1;
        st->kernel_launches += split ? (uint64_t)split_launches + (split_mode == 2 ? 0u : (uint64_t)(n_levels - 1))
                                     : (uint64_t)n_levels * (pretrace ? (defer_shadows ? 3u : 2u) : 1u);
        st->level_kernel_launches += (uint64_t)n_levels;
        if (warp) st->warp_kernel_launches += (uint64_t)(n_levels - 1);
        st->peak_ray_records = std::max<uint64_t>(st->peak_ray_records, peak_r);
        st->peak_fan_records = std::max<uint64_t>(st->peak_fan_records, peak_f);
        for (int L = 0; L < n_levels; ++L) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, sl.events[L], sl.events[L + 1]);
            st->level_ms[L] += ms;
            st->level_kernel_ms += ms;
        }
        // records written by level L-1 and read by level L: 48 B each way
        for (int L = 1; L <= n_levels; ++L)
            for (int k = 0; k < 1 + SP_MAX_FAN_CLASSES; ++k) st->queue_bytes += 2ull * 48ull * counts[(size_t)L * ncl + k];
        for (int L = 0; L < SP_MAX_LEVELS && L < SP_MAX_DEPTH_LEVELS; ++L) {
            st->rays_per_depth[L] += ds.rays[L];
            st->rays_total += ds.rays[L];
        }
        st->shadow_rays += ds.shadow_rays;
    }
    if (getenv("SIGHTPY_PHASE_TIMING")) {             // only meaningful for -DSP_PHASE_TIMING builds of sp_kernels.cu
        double tot = 0;
        for (int k = 0; k < 6; ++k) tot += (double)ds.phase_cycles[k];
        if (tot > 0)
            fprintf(stderr, "[sightpy-b200] warp-cycles: generate %.1f %%, intersect %.1f %%, park+count %.1f %%, wait A %.1f %%, "
                            "shade %.1f %%, wait C %.1f %%\n", 100.0 * ds.phase_cycles[0] / tot, 100.0 * ds.phase_cycles[1] / tot,
                    100.0 * ds.phase_cycles[2] / tot, 100.0 * ds.phase_cycles[3] / tot, 100.0 * ds.phase_cycles[4] / tot,
                    100.0 * ds.phase_cycles[5] / tot);
    }
    return 0;
}


// One chunk, synchronously (caller rays, single passes): enqueue, wait.
static int run_chunk(sp_scene* s, const ChunkJob& job, sp_stats* st, bool& overflow) {
    overflow = false;
    bool dummy = false;
    if (int rc = finish_chunk(s, 0, nullptr, dummy)) return rc;          // (nothing is ever left in flight between calls)
    if (int rc = enqueue_chunk(s, job, 0, nullptr, nullptr, 0u)) return rc;
    return finish_chunk(s, 0, st, overflow);
}

static int begin_call(sp_scene* s, uint64_t seed, sp_stats* st, const char* what, uint64_t primaries) {
    if (!s) return fail("%s: null scene", what);
    if (!s->committed) return fail("%s: scene not committed (sp_scene_commit)", what);
    s->d.seed_lo = (uint32_t)(seed & 0xFFFFFFFFull);
    s->d.seed_hi = (uint32_t)(seed >> 32);
    for (uint32_t r = 0; r < 10; ++r) {
        s->d.philox_keys[2 * r] = s->d.seed_lo + r * 0x9E3779B9u;
        s->d.philox_keys[2 * r + 1] = s->d.seed_hi + r * 0xBB67AE85u;
    }
    if (st) memset(st, 0, sizeof *st);
    s->call_primaries = primaries;
    return ensure_queues(s, primaries);
}

static int end_call(sp_scene* s, sp_stats* st, cudaEvent_t t0, cudaEvent_t t1) {
    CUDA_TRY(cudaEventRecord(t1, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (st) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        st->device_ms = ms;
    }
    return 0;
}

// Primaries of the next chunk given the queue occupancy seen so far.  Before anything is known the first chunk is
// one sample of the whole region when that is at most 16 Mi primaries (representative of the frame, unlike its
// first rows), else a 64 Ki-primary probe.
static uint32_t pick_chunk(sp_scene* s, uint64_t region) {
    int64_t p = default_chunk(s);
    if (s->use_ray == 0.0 && s->use_fan == 0.0)
        p = std::min<int64_t>(p, region <= ((uint64_t)16 << 20) ? (int64_t)std::max<uint64_t>(region, 1) : (int64_t)1 << 16);
    else if (s->opt_chunk == 0)
        p *= 8;            // scenes that queue little (Whitted trees: 1-3 records per primary) take far larger chunks in the same queues
    const double slack = 1.3;
    if (s->use_ray > 0.0) p = std::min<int64_t>(p, (int64_t)(s->ray_cap / (s->use_ray * slack)));
    if (s->use_fan > 0.0) p = std::min<int64_t>(p, (int64_t)(s->fan_cap / (s->use_fan * slack)));
    if (s->chunk_limit > 0) p = std::min<int64_t>(p, s->chunk_limit);
    return (uint32_t)std::min<int64_t>(std::max<int64_t>(p, 1024), (int64_t)0x7FFFFFFF);
}

// A chunk overflowed a queue: what to do next.  Returns non-zero (with the overflow message) when the chunk cannot
// get any smaller.
static int shrink_after_overflow(sp_scene* s, uint32_t n_items, sp_stats* st) {
    if (st) st->chunk_retries += 1;
    if (n_items <= 1024u || s->opt_chunk_fixed) {
        const std::string msg = g_error;
        return fail("%s: raise ray_queue_capacity / fan_queue_capacity or lower chunk_primaries", msg.c_str());
    }
    s->chunk_limit = std::max<int64_t>(n_items / 2, 1024);
    return 0;
}

// Render samples [sample_begin, sample_end) of a region of n_region "pixels" starting at first_pix — plain pixel
// indices, or (tiles != nullptr) texels of a tile list — into the accumulation buffer, chunk by chunk.
static int render_chunks(sp_scene* s, uint32_t first_pix, uint32_t n_region, const uint32_t* tiles, uint32_t tile_shift,
                         uint32_t tiles_x, int sample_begin, int sample_end, int clear, sp_stats* st, const char* what) {
    ScopedEvent ev0, ev1;
    if (!ev0.e || !ev1.e) return fail("%s: cudaEventCreate failed", what);
    cudaEvent_t t0 = ev0.e, t1 = ev1.e;
    // Chunks add their radiance to a scratch frame that is folded into the accumulation buffer once the chunk is
    // known to be complete (on the device: sp_fold_kernel looks at the chunk's overflow flag), so that a chunk whose
    // queues overflowed can be rendered again in smaller pieces.  Two chunks are in flight: chunk k + 1 is enqueued
    // before the host waits for chunk k's counters, each with a scratch frame and counters of its own.
    for (auto& sl : s->slot)
        if (sl.scratch.n != s->accum.n) {
            CUDA_TRY(sl.scratch.alloc(s->accum.n));
            CUDA_TRY(cudaMemsetAsync(sl.scratch.p, 0, sl.scratch.n * sizeof(float4), s->stream));
        }
    CUDA_TRY(cudaEventRecord(t0, s->stream));
    if (clear) CUDA_TRY(cudaMemsetAsync(s->accum.p, 0, s->accum.n * sizeof(float4), s->stream));
    // what is left to render: (pixel range inside the region) x (sample range); a chunk is carved off the front
    struct Todo { uint32_t pix0, npix, s0, s1; };
    std::deque<Todo> todo;
    if (n_region > 0 && sample_end > sample_begin) todo.push_back({0u, n_region, (uint32_t)sample_begin, (uint32_t)sample_end});
    int rc = 0, next_slot = 0;
    auto finish = [&](int k) -> int {
        ChunkSlot& sl = s->slot[k];
        if (!sl.busy) return 0;
        bool overflow = false;
        const uint32_t pix0 = sl.region_pix, npix = sl.job.n_pix, s0 = sl.job.sample_begin, ns = sl.region_ns, n_items = sl.job.n_items;
        if (int r = finish_chunk(s, k, st, overflow)) return r;
        if (overflow) {                                         // nothing of it was folded: again, in smaller pieces
            todo.push_front({pix0, npix, s0, s0 + ns});
            return shrink_after_overflow(s, n_items, st);
        }
        return 0;
    };
    while (rc == 0 && !todo.empty()) {
        const uint32_t P = pick_chunk(s, n_region);
        ChunkSlot& sl = s->slot[next_slot];
        ChunkJob job{};
        job.source = SP_SRC_CAMERA; job.run = SP_RUN_FULL; job.accum = sl.scratch.p;
        job.tiles = tiles; job.tile_shift = tile_shift; job.tiles_x = tiles_x;
        Todo r = todo.front();
        todo.pop_front();
        uint32_t ns = 1;
        if (P >= r.npix) {                                       // the whole pixel range x several samples
            ns = std::max<uint32_t>(std::min<uint32_t>(P / r.npix, r.s1 - r.s0), 1u);
            job.pix_begin = first_pix + r.pix0; job.n_pix = r.npix; job.sample_begin = r.s0; job.n_items = ns * r.npix;
            if (r.s0 + ns < r.s1) todo.push_front({r.pix0, r.npix, r.s0 + ns, r.s1});
        } else {                                                 // a slice of the pixel range, one sample
            const uint32_t n = std::min<uint32_t>(P, r.npix);
            job.pix_begin = first_pix + r.pix0; job.n_pix = n; job.sample_begin = r.s0; job.n_items = n;
            if (r.s0 + 1 < r.s1) todo.push_front({r.pix0, r.npix, r.s0 + 1, r.s1});
            if (n < r.npix) todo.push_front({r.pix0 + n, r.npix - n, r.s0, r.s0 + 1});
        }
        // the part of the frame the chunk can touch: its pixel range, or (tiles) anything
        float4* const acc_lo = tiles ? s->accum.p : s->accum.p + job.pix_begin;
        float4* const scr_lo = tiles ? sl.scratch.p : sl.scratch.p + job.pix_begin;
        const uint32_t n_touch = tiles ? (uint32_t)s->accum.n : job.n_pix;
        // nothing is known about the scene's queue occupancy yet: this chunk is the probe, wait for it
        const bool probe = s->use_ray == 0.0 && s->use_fan == 0.0;
        rc = enqueue_chunk(s, job, next_slot, acc_lo, scr_lo, n_touch);
        if (rc) break;
        sl.region_pix = r.pix0; sl.region_ns = ns;
        if (probe) rc = finish(next_slot);
        else { rc = finish(next_slot ^ 1); next_slot ^= 1; }      // wait for the chunk before this one while this one runs
        if (rc == 0 && todo.empty()) {                            // the last chunks: nothing left to overlap them with
            rc = finish(0);
            if (rc == 0) rc = finish(1);
        }
    }
    for (int k = 0; k < 2; ++k) {                                // (an error above may have left a chunk in flight)
        bool ov = false;
        if (s->slot[k].busy) { const int r2 = finish_chunk(s, k, nullptr, ov); if (rc == 0) rc = r2; }
    }
    int rc2 = end_call(s, st, t0, t1);
    return rc ? rc : rc2;
}

int sp_render_region(sp_scene* s, int64_t pix_begin, int64_t pix_end, int sample_begin, int sample_end, uint64_t seed,
                     int clear, sp_stats* st) {
    SP_ENTER(s ? s->device : (g_device >= 0 ? g_device : 0));
    if (s && s->committed && !s->has_camera) return fail("sp_render_region: the scene has no camera");
    const uint64_t region = (s && s->committed && pix_end > pix_begin) ? (uint64_t)(pix_end - pix_begin) : 0;
    int rc = begin_call(s, seed, st, "sp_render_region", (uint64_t)std::max(sample_end - sample_begin, 0) * region);
    if (rc) return rc;
    const uint32_t n_pix_total = (uint32_t)s->d.cam.W * (uint32_t)s->d.cam.H;
    if (sample_begin < 0 || sample_end < sample_begin) return fail("sp_render_region: invalid sample range");
    if (pix_begin < 0 || pix_end < pix_begin || pix_end > (int64_t)n_pix_total) return fail("sp_render_region: invalid pixel range");
    return render_chunks(s, (uint32_t)pix_begin, (uint32_t)(pix_end - pix_begin), nullptr, 0u, 0u, sample_begin, sample_end, clear, st,
                         "sp_render_region");
}

int sp_render_tiles(sp_scene* s, const int32_t* tile_ids, int n_tiles, int tile_size, int sample_begin, int sample_end,
                    uint64_t seed, int clear, sp_stats* st) {
    SP_ENTER(s ? s->device : (g_device >= 0 ? g_device : 0));
    if (s && s->committed && !s->has_camera) return fail("sp_render_tiles: the scene has no camera");
    if (n_tiles < 0 || (n_tiles > 0 && !tile_ids)) return fail("sp_render_tiles: invalid tile list");
    int shift = 0;
    while ((1 << shift) < tile_size) ++shift;
    if (tile_size < 1 || tile_size > 1024 || (1 << shift) != tile_size) return fail("sp_render_tiles: tile_size must be a power of two in 1..1024");
    const uint64_t region = (uint64_t)std::max(n_tiles, 0) * (uint64_t)tile_size * (uint64_t)tile_size;
    if (region > 0x7FFFFFFFull) return fail("sp_render_tiles: too many tiles");
    int rc = begin_call(s, seed, st, "sp_render_tiles", (uint64_t)std::max(sample_end - sample_begin, 0) * region);
    if (rc) return rc;
    if (sample_begin < 0 || sample_end < sample_begin) return fail("sp_render_tiles: invalid sample range");
    const int tiles_x = (s->d.cam.W + tile_size - 1) / tile_size, tiles_y = (s->d.cam.H + tile_size - 1) / tile_size;
    std::vector<uint32_t> ids((size_t)n_tiles);
    for (int i = 0; i < n_tiles; ++i) {
        if (tile_ids[i] < 0 || tile_ids[i] >= tiles_x * tiles_y) return fail("sp_render_tiles: tile %d out of range (frame has %d x %d tiles)", tile_ids[i], tiles_x, tiles_y);
        ids[(size_t)i] = (uint32_t)tile_ids[i];
    }
    CUDA_TRY(s->d_tiles.upload(ids));
    return render_chunks(s, 0u, (uint32_t)region, s->d_tiles.p, (uint32_t)shift, (uint32_t)tiles_x, sample_begin, sample_end, clear, st,
                         "sp_render_tiles");
}

int sp_render_samples(sp_scene* s, int sample_begin, int sample_end, uint64_t seed, int clear, sp_stats* st) {
    if (!s || !s->committed) return fail("sp_render_samples: scene not committed (sp_scene_commit)");
    if (!s->has_camera) return fail("sp_render_samples: the scene has no camera");
    return sp_render_region(s, 0, (int64_t)s->accum.n, sample_begin, sample_end, seed, clear, st);
}

void* sp_accum_device_ptr(sp_scene* s) { return s ? (void*)s->accum.p : nullptr; }
uint64_t sp_accum_bytes(sp_scene* s) { return s ? (uint64_t)(s->accum.n * sizeof(float4)) : 0; }

int sp_resolve(sp_scene* s, int spp_total, float* out_linear, uint8_t* out_srgb8) {
    SP_ENTER(s ? s->device : (g_device >= 0 ? g_device : 0));
    if (!s || !s->committed || !s->has_camera) return fail("sp_resolve: scene not committed or has no camera");
    if (spp_total < 1) return fail("sp_resolve: spp_total must be >= 1");
    const size_t n = s->accum.n;
    if (s->d_lin.n != 3 * n) CUDA_TRY(s->d_lin.alloc(3 * n));
    if (s->d_u8.n != 3 * n) CUDA_TRY(s->d_u8.alloc(3 * n));
    ResolveArgs a;
    a.accum = s->accum.p; a.n_pix = (uint32_t)n; a.spp = (double)spp_total; a.out_linear = s->d_lin.p; a.out_srgb8 = s->d_u8.p;
    CUDA_TRY(sp_launch_resolve(a, s->stream));
    if (out_linear) CUDA_TRY(cudaMemcpyAsync(out_linear, s->d_lin.p, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    if (out_srgb8) CUDA_TRY(cudaMemcpyAsync(out_srgb8, s->d_u8.p, 3 * n, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}

int sp_scene_set_stream(sp_scene* s, void* cuda_stream, int use_it) {
    SP_ENTER(s ? s->device : (g_device >= 0 ? g_device : 0));
    if (!s) return fail("sp_scene_set_stream: null scene");
    s->user_stream_set = use_it != 0;
    s->user_stream = (cudaStream_t)cuda_stream;
    if (s->committed) {
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        s->stream = s->user_stream_set ? s->user_stream : s->own_stream;
    }
    return 0;
}

int sp_render(sp_scene* s, int spp, uint64_t seed, float* out_linear, uint8_t* out_srgb8, sp_stats* st) {
    if (spp < 1) return fail("sp_render: samples_per_pixel must be >= 1");
    int rc = sp_render_samples(s, 0, spp, seed, 1, st);
    if (rc) return rc;
    if (st) st->kernel_launches += 1;
    return sp_resolve(s, spp, out_linear, out_srgb8);
}

int sp_trace(sp_scene* s, const float* origins, const float* dirs, int n, uint64_t seed, float* out_rgb,
             int32_t* out_hit_id, float* out_t, sp_stats* st) {
    SP_ENTER(s ? s->device : (g_device >= 0 ? g_device : 0));
    int rc = begin_call(s, seed, st, "sp_trace", (uint64_t)std::max(n, 0));
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!origins || !dirs))) return fail("sp_trace: invalid arguments");
    if (n == 0) return 0;
    DevBuf<float> d_o, d_d, d_t;
    DevBuf<float4> d_acc;
    DevBuf<int32_t> d_hit;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    auto cleanup = [&]() {
        d_o.release(); d_d.release(); d_t.release(); d_acc.release(); d_hit.release();
        if (t0) cudaEventDestroy(t0);
        if (t1) cudaEventDestroy(t1);
    };
#define TRY_OR_CLEAN(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cleanup(); return fail("%s: %s", #expr, cudaGetErrorString(e__)); } } while (0)
    TRY_OR_CLEAN(d_o.alloc(3 * (size_t)n)); TRY_OR_CLEAN(d_d.alloc(3 * (size_t)n));
    TRY_OR_CLEAN(d_t.alloc((size_t)n)); TRY_OR_CLEAN(d_hit.alloc((size_t)n)); TRY_OR_CLEAN(d_acc.alloc((size_t)n));
    TRY_OR_CLEAN(cudaMemcpyAsync(d_o.p, origins, 3 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    TRY_OR_CLEAN(cudaMemcpyAsync(d_d.p, dirs, 3 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    TRY_OR_CLEAN(cudaMemsetAsync(d_acc.p, 0, (size_t)n * sizeof(float4), s->stream));
    TRY_OR_CLEAN(cudaEventCreate(&t0)); TRY_OR_CLEAN(cudaEventCreate(&t1));
    TRY_OR_CLEAN(cudaEventRecord(t0, s->stream));
    for (uint32_t base = 0; base < (uint32_t)n && rc == 0;) {
        const uint32_t P = std::min<uint32_t>(pick_chunk(s, (uint64_t)n), (uint32_t)n - base);
        ChunkJob job{};
        job.source = SP_SRC_USER; job.run = SP_RUN_FULL; job.accum = d_acc.p;
        job.user_base = base; job.n_items = P; job.user_o = d_o.p; job.user_d = d_d.p;
        job.out_hit = d_hit.p; job.out_t = d_t.p;
        bool overflow = false;
        rc = run_chunk(s, job, st, overflow);
        if (rc) break;
        if (overflow) {                                          // a ray's radiance lands in its own element: clear the chunk's and go again
            TRY_OR_CLEAN(cudaMemsetAsync(d_acc.p + base, 0, (size_t)P * sizeof(float4), s->stream));
            rc = shrink_after_overflow(s, P, st);
            continue;
        }
        base += P;
    }
    int rc2 = end_call(s, st, t0, t1);
    if (rc == 0) rc = rc2;
    if (rc == 0) {
        if (out_rgb) {
            std::vector<float4> acc((size_t)n);
            TRY_OR_CLEAN(cudaMemcpy(acc.data(), d_acc.p, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost));
            for (int i = 0; i < n; ++i) { out_rgb[3 * (size_t)i] = acc[i].x; out_rgb[3 * (size_t)i + 1] = acc[i].y; out_rgb[3 * (size_t)i + 2] = acc[i].z; }
        }
        if (out_hit_id) TRY_OR_CLEAN(cudaMemcpy(out_hit_id, d_hit.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost));
        if (out_t) TRY_OR_CLEAN(cudaMemcpy(out_t, d_t.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
    }
    cleanup();
    return rc;
}

static int primary_pass(sp_scene* s, int run, int sample, uint64_t seed, float* out_o, float* out_d, float* out_t,
                        int32_t* out_hit, float* out_n, const char* what) {
    SP_ENTER(s ? s->device : (g_device >= 0 ? g_device : 0));
    int rc = begin_call(s, seed, nullptr, what, s && s->committed ? s->accum.n : 0);
    if (rc) return rc;
    if (!s->has_camera) return fail("%s: the scene has no camera", what);
    const size_t n = s->accum.n;
    DevBuf<float> d_o, d_d, d_t, d_n;
    DevBuf<int32_t> d_hit;
    auto cleanup = [&]() { d_o.release(); d_d.release(); d_t.release(); d_n.release(); d_hit.release(); };
    if (run == SP_RUN_DUMP_RAYS) { TRY_OR_CLEAN(d_o.alloc(3 * n)); TRY_OR_CLEAN(d_d.alloc(3 * n)); }
    else {
        if (out_t) TRY_OR_CLEAN(d_t.alloc(n));
        if (out_hit) TRY_OR_CLEAN(d_hit.alloc(n));
        if (out_n) TRY_OR_CLEAN(d_n.alloc(3 * n));
    }
    ChunkJob job{};
    job.source = SP_SRC_CAMERA; job.run = run; job.accum = s->accum.p;
    job.pix_begin = 0; job.n_pix = (uint32_t)n; job.sample_begin = (uint32_t)sample; job.n_items = (uint32_t)n;
    job.out_o = d_o.p; job.out_d = d_d.p; job.out_t = d_t.p; job.out_hit = d_hit.p; job.out_n = d_n.p;
    bool overflow = false;                                    // a level-0-only pass queues nothing
    rc = run_chunk(s, job, nullptr, overflow);
    if (rc == 0 && run == SP_RUN_DUMP_RAYS) {
        TRY_OR_CLEAN(cudaMemcpy(out_o, d_o.p, 3 * n * sizeof(float), cudaMemcpyDeviceToHost));
        TRY_OR_CLEAN(cudaMemcpy(out_d, d_d.p, 3 * n * sizeof(float), cudaMemcpyDeviceToHost));
    } else if (rc == 0) {
        if (out_t) TRY_OR_CLEAN(cudaMemcpy(out_t, d_t.p, n * sizeof(float), cudaMemcpyDeviceToHost));
        if (out_hit) TRY_OR_CLEAN(cudaMemcpy(out_hit, d_hit.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
        if (out_n) TRY_OR_CLEAN(cudaMemcpy(out_n, d_n.p, 3 * n * sizeof(float), cudaMemcpyDeviceToHost));
    }
    cleanup();
    return rc;
}

int sp_camera_rays(sp_scene* s, int sample, uint64_t seed, float* out_origins, float* out_dirs) {
    if (!out_origins || !out_dirs || sample < 0) return fail("sp_camera_rays: invalid arguments");
    return primary_pass(s, SP_RUN_DUMP_RAYS, sample, seed, out_origins, out_dirs, nullptr, nullptr, nullptr, "sp_camera_rays");
}

int sp_distances(sp_scene* s, uint64_t seed, float* out_t) {
    if (!out_t) return fail("sp_distances: invalid arguments");
    return primary_pass(s, SP_RUN_DISTANCES, 0, seed, nullptr, nullptr, out_t, nullptr, nullptr, "sp_distances");
}

int sp_aovs(sp_scene* s, int sample, uint64_t seed, int32_t* out_hit_id, float* out_t, float* out_normal) {
    if ((!out_hit_id && !out_t && !out_normal) || sample < 0) return fail("sp_aovs: invalid arguments");
    return primary_pass(s, SP_RUN_DISTANCES, sample, seed, nullptr, nullptr, out_t, out_hit_id, out_normal, "sp_aovs");
}

int sp_scene_read_texture(sp_scene* s, int tex_id, uint8_t* out_rgb) {
    SP_ENTER(s ? s->device : (g_device >= 0 ? g_device : 0));
    if (!s || !s->committed || !out_rgb) return fail("sp_scene_read_texture: scene not committed or null output");
    if (tex_id < 0 || tex_id >= (int)s->textures.size()) return fail("sp_scene_read_texture: texture %d out of range", tex_id);
    const HostTexture& ht = s->textures[(size_t)tex_id];
    const size_t n = (size_t)ht.H * ht.W;
    DTexture desc;
    CUDA_TRY(cudaMemcpy(&desc, s->d_texdesc.p + tex_id, sizeof desc, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> texels(n);
    CUDA_TRY(cudaMemcpy(texels.data(), desc.texels, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; ++i) {
        out_rgb[3 * i] = (uint8_t)(texels[i] & 255u); out_rgb[3 * i + 1] = (uint8_t)((texels[i] >> 8) & 255u);
        out_rgb[3 * i + 2] = (uint8_t)((texels[i] >> 16) & 255u);
    }
    return 0;
}

int sp_set_option(sp_scene* s, const char* name, int64_t value) {
    if (!s || !name) return fail("sp_set_option: invalid arguments");
    if (value < 0) return fail("sp_set_option: %s must be >= 0", name);
    if (!strcmp(name, "ray_queue_capacity")) s->opt_ray_cap = value;
    else if (!strcmp(name, "fan_queue_capacity")) s->opt_fan_cap = value;
    else if (!strcmp(name, "chunk_primaries")) s->opt_chunk = value;
    else if (!strcmp(name, "fixed_chunks")) s->opt_chunk_fixed = value;
    else if (!strcmp(name, "pretrace")) s->opt_pretrace = value;
    else if (!strcmp(name, "split_kernels")) {
        if (value > 2) return fail("sp_set_option: split_kernels must be 0 (off), 1 (level 0) or 2 (every level)");
        s->opt_split = value; s->d.use_split = value ? 1 : 0;
    }
    else if (!strcmp(name, "max_levels")) s->opt_max_levels = value;
    else if (!strcmp(name, "bvh")) {
        s->opt_bvh = value;
        if (s->committed) return sp_scene_commit(s);             // the geometry tables depend on it
    }
    else if (!strcmp(name, "warp_kernel")) {
        s->opt_warp = value;
        if (s->committed) {
            s->d.use_warp_kernel = value ? 1 : 0;
            if (int rc = pick_kernels(s)) return rc;
        }
    }
    else return fail("sp_set_option: unknown option '%s'", name);
    s->use_ray = s->use_fan = 0.0;
    s->chunk_limit = 0;
    return 0;
}

// ---- several GPUs of one node, driven by the library itself (the reference's process pool, scene.py:98-116) -----------
// `scenes` are n committed replicas of one scene on n different devices (sp_scene_create_on).  One host thread per
// device renders that device's shard of the frame — a contiguous sample range, or (shard_mode 1, and whenever spp < n)
// the interleaved 64x64 tiles r, r + n, ... — then the first scene's device adds the other frames to its own, reading
// them over NVLink (peer access, sp_init_devices; a staged peer copy otherwise), and resolves.
int sp_render_group(sp_scene** scenes, int n, int spp, uint64_t seed, int shard_mode, float* out_linear, uint8_t* out_srgb8,
                    sp_stats* st) {
    if (!scenes || n < 1) return fail("sp_render_group: need at least one scene");
    if (spp < 1) return fail("sp_render_group: samples_per_pixel must be >= 1");
    for (int r = 0; r < n; ++r) {
        if (!scenes[r] || !scenes[r]->committed || !scenes[r]->has_camera) return fail("sp_render_group: scene %d is not committed with a camera", r);
        if (scenes[r]->accum.n != scenes[0]->accum.n) return fail("sp_render_group: scene %d has a different frame size", r);
        for (int q = 0; q < r; ++q)
            if (scenes[q]->device == scenes[r]->device) return fail("sp_render_group: scenes %d and %d share device %d", q, r, scenes[r]->device);
    }
    const bool tiles = shard_mode == 1 || spp < n;
    const int W = scenes[0]->d.cam.W, H = scenes[0]->d.cam.H, T = 64;
    const int n_tiles = ((W + T - 1) / T) * ((H + T - 1) / T);
    std::vector<int> rcs((size_t)n, 0);
    std::vector<std::string> errs((size_t)n);
    std::vector<sp_stats> sts((size_t)n);
    auto work = [&](int r) {
        if (tiles) {
            std::vector<int32_t> ids;
            for (int t = r; t < n_tiles; t += n) ids.push_back(t);
            rcs[(size_t)r] = sp_render_tiles(scenes[r], ids.data(), (int)ids.size(), T, 0, spp, seed, 1, &sts[(size_t)r]);
        } else {
            const int base = spp / n, extra = spp % n;
            const int begin = r * base + std::min(r, extra), end = begin + base + (r < extra ? 1 : 0);
            rcs[(size_t)r] = sp_render_samples(scenes[r], begin, end, seed, 1, &sts[(size_t)r]);
        }
        if (rcs[(size_t)r]) errs[(size_t)r] = g_error;             // the message is thread-local
    };
    std::vector<std::thread> threads;
    for (int r = 1; r < n; ++r) threads.emplace_back(work, r);
    work(0);
    for (auto& t : threads) t.join();
    for (int r = 0; r < n; ++r)
        if (rcs[(size_t)r]) return fail("sp_render_group: device %d: %s", scenes[r]->device, errs[(size_t)r].c_str());
    sp_scene* s0 = scenes[0];
    {
        SP_ENTER(s0->device);
        for (int r = 1; r < n; ++r) {
            int can = 0;
            cudaDeviceCanAccessPeer(&can, s0->device, scenes[r]->device);
            const float4* src = scenes[r]->accum.p;
            if (can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(scenes[r]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
                cudaGetLastError();
            }
            if (!can) {                                          // no direct access: stage the peer's frame in the scratch frame
                CUDA_TRY(cudaMemcpyPeerAsync(s0->slot[0].scratch.p, s0->device, scenes[r]->accum.p, scenes[r]->device,
                                             s0->accum.n * sizeof(float4), s0->stream));
                src = s0->slot[0].scratch.p;
            }
            CUDA_TRY(sp_launch_add(s0->accum.p, src, (uint32_t)s0->accum.n, s0->stream));
            if (!can) CUDA_TRY(cudaMemsetAsync(s0->slot[0].scratch.p, 0, s0->accum.n * sizeof(float4), s0->stream));
        }
        CUDA_TRY(cudaStreamSynchronize(s0->stream));
    }
    if (st) {
        *st = sts[0];
        for (int r = 1; r < n; ++r) {
            const sp_stats& o = sts[(size_t)r];
            st->rays_total += o.rays_total; st->shadow_rays += o.shadow_rays; st->kernel_launches += o.kernel_launches;
            st->chunks += o.chunks; st->level_kernel_launches += o.level_kernel_launches; st->queue_bytes += o.queue_bytes;
            st->warp_kernel_launches += o.warp_kernel_launches; st->chunk_retries += o.chunk_retries;
            st->device_ms = std::max(st->device_ms, o.device_ms);
            st->level_kernel_ms = std::max(st->level_kernel_ms, o.level_kernel_ms);
            st->peak_ray_records = std::max(st->peak_ray_records, o.peak_ray_records);
            st->peak_fan_records = std::max(st->peak_fan_records, o.peak_fan_records);
            for (int L = 0; L < SP_MAX_DEPTH_LEVELS; ++L) { st->rays_per_depth[L] += o.rays_per_depth[L]; st->level_ms[L] = std::max(st->level_ms[L], o.level_ms[L]); }
        }
        st->kernel_launches += (uint64_t)n;                      // the peer adds and the resolve
    }
    return sp_resolve(s0, spp, out_linear, out_srgb8);
}

int sp_measure_peaks(double* fp32_tflops, double* copy_gbs) {
    if (g_device < 0) return fail("sp_measure_peaks: call sp_init first");
    SP_ENTER(g_device);
    double a = 0.0, b = 0.0;
    CUDA_TRY(sp_bench_ffma(&a, nullptr));
    CUDA_TRY(sp_bench_copy(&b, nullptr));
    if (fp32_tflops) *fp32_tflops = a;
    if (copy_gbs) *copy_gbs = b;
    return 0;
}

}  // extern "C"
