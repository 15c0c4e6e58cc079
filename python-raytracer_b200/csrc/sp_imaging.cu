// Sky-box blur on the device: restates blur_skybox (sightpy/backgrounds/util/blur_background.py:17-132)
// bit for bit on RGBA8 texels.
//
// The reference blurs a cross-layout cube map face by face: each face is pasted, with its four edge neighbours
// rotated into place, onto a 3N x 3N canvas (blur_background.py:40-118) so that the blur bleeds correctly across
// cube edges; the canvas is quantised to bytes with (255 * x).astype(uint8) on x = byte / 256 (to_image, :6-10:
// byte b becomes max(b - 1, 0)), filtered with Pillow's ImageFilter.GaussianBlur and the centre tile is kept.
// Pillow's GaussianBlur is three passes of a fixed-point box filter per axis (libImaging/BoxBlur.c): the box radius
// follows from sigma (Gwosdek et al., "Theoretical foundations of Gaussian convolution by extended box filtering"),
// every pass re-quantises to bytes with   (acc * ww + (far_left + far_right) * fw + 2^23) >> 24   in uint32, where
// acc is the sum over the integer window with replicated edges, ww = (uint32)(2^24 / (2 r + 1)) for the fractional
// radius r (float arithmetic) and fw the weight of the two pixels just outside the window.  A pass along rows followed
// by a transposition is a pass along columns, so no transposition is needed here.  tests/test_gpu_parity.py compares
// the result with the reference's own output (sha256 recorded by tests/golden/make_golden.py).
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>

// (source face, quarter turns counter-clockwise) of the five canvas tiles left, centre, right, below, above, per
// target face; faces numbered left 0, front 1, right 2, back 3, top 4, bottom 5 (blur_background.py:40-118)
struct CanvasTile { int8_t face, turns; };
__constant__ CanvasTile c_canvas[6][5] = {
    /* left   */ {{3, 0}, {0, 0}, {1, 0}, {5, -1}, {4, 1}},
    /* front  */ {{0, 0}, {1, 0}, {2, 0}, {5, 0}, {4, 0}},
    /* right  */ {{1, 0}, {2, 0}, {3, 0}, {5, 1}, {4, -1}},
    /* back   */ {{2, 0}, {3, 0}, {0, 0}, {5, 2}, {4, 2}},
    /* top    */ {{0, -1}, {4, 0}, {2, 1}, {1, 0}, {3, 2}},
    /* bottom */ {{0, 1}, {5, 0}, {2, -1}, {3, 2}, {1, 0}},
};
// (row block, column block) of each face inside the 3 x 4 cross image
__constant__ int2 c_slot[6] = {{1, 0}, {1, 1}, {1, 2}, {1, 3}, {0, 1}, {2, 1}};

__device__ __forceinline__ uint32_t requant(uint32_t texel) {       // byte b -> (uint8)(255 * (b / 256)) = max(b - 1, 0)
    uint32_t r = texel & 255u, g = (texel >> 8) & 255u, b = (texel >> 16) & 255u;
    r = r ? r - 1u : 0u; g = g ? g - 1u : 0u; b = b ? b - 1u : 0u;
    return r | (g << 8) | (b << 16);
}

// canvas[y][x] of target face f: zero outside the five tiles
__global__ void __launch_bounds__(256) sp_blur_canvas_kernel(const uint32_t* __restrict__ cross, uint32_t* __restrict__ canvas,
                                                             int N, int W, int f) {
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= 3 * N) return;
    const int by = y / N, bx = x / N;                              // canvas block
    int tile = -1;
    if (by == 1) tile = bx;                                        // left, centre, right
    else if (bx == 1) tile = by == 2 ? 3 : 4;                      // below, above
    uint32_t v = 0u;
    if (tile >= 0) {
        const CanvasTile t = c_canvas[f][tile];
        const int i = y - by * N, j = x - bx * N;                  // position inside the rotated tile
        int si = i, sj = j;                                        // np.rot90(face, k)[i][j]
        const int k = ((int)t.turns % 4 + 4) % 4;
        if (k == 1) { si = j; sj = N - 1 - i; }
        else if (k == 2) { si = N - 1 - i; sj = N - 1 - j; }
        else if (k == 3) { si = N - 1 - j; sj = i; }
        const int2 slot = c_slot[t.face];
        v = requant(cross[(size_t)(slot.x * N + si) * W + (slot.y * N + sj)]);
    }
    canvas[(size_t)y * (3 * N) + x] = v;
}

// one box-filter pass along rows (vertical == 0) or columns of a size x size canvas
__global__ void __launch_bounds__(256) sp_blur_box_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int size,
                                                          int radius, uint32_t ww, uint32_t fw, int vertical) {
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= size) return;
    const int pos = vertical ? y : x;
    const size_t line = vertical ? (size_t)x : (size_t)y * size, stride = vertical ? (size_t)size : 1;
    auto at = [&](int p) { p = p < 0 ? 0 : (p > size - 1 ? size - 1 : p); return in[line + (size_t)p * stride]; };
    uint32_t ar = 0, ag = 0, ab = 0;
    for (int k = -radius; k <= radius; ++k) {
        const uint32_t t = at(pos + k);
        ar += t & 255u; ag += (t >> 8) & 255u; ab += (t >> 16) & 255u;
    }
    const uint32_t l = at(pos - radius - 1), r = at(pos + radius + 1);
    const uint32_t br = ar * ww + ((l & 255u) + (r & 255u)) * fw;
    const uint32_t bg = ag * ww + (((l >> 8) & 255u) + ((r >> 8) & 255u)) * fw;
    const uint32_t bb = ab * ww + (((l >> 16) & 255u) + ((r >> 16) & 255u)) * fw;
    out[(size_t)y * size + x] = ((br + (1u << 23)) >> 24) | (((bg + (1u << 23)) >> 24) << 8) | (((bb + (1u << 23)) >> 24) << 16);
}

__global__ void __launch_bounds__(256) sp_blur_extract_kernel(const uint32_t* __restrict__ canvas, uint32_t* __restrict__ out, int N,
                                                              int W, int f) {
    const int j = blockIdx.x * 256 + threadIdx.x, i = blockIdx.y;
    if (j >= N) return;
    const int2 slot = c_slot[f];
    out[(size_t)(slot.x * N + i) * W + (slot.y * N + j)] = canvas[(size_t)(N + i) * (3 * N) + (N + j)];
}

// Pillow's _gaussian_blur_radius (libImaging/BoxBlur.c), float arithmetic as there
static float box_radius(float radius, int passes) {
    const float sigma2 = radius * radius / passes;
    const float L = sqrtf(12.0f * sigma2 + 1.0f);
    const float l = floorf((L - 1.0f) / 2.0f);
    float a = (2 * l + 1) * (l * (l + 1) - 3 * sigma2);
    a /= 6 * (sigma2 - (l + 1) * (l + 1));
    return l + a;
}

// cross: H x W packed texels (r | g << 8 | b << 16) of a cross-layout cube map, N = H / 3; out: the blurred cross
// (texels outside the six faces are zero, as in the reference); tmp0 / tmp1: two (3N)^2 scratch canvases.
cudaError_t sp_blur_cube_cross(const uint32_t* cross, uint32_t* out, uint32_t* tmp0, uint32_t* tmp1, int H, int W, float blur,
                               cudaStream_t st) {
    const int N = H / 3, size = 3 * N, passes = 3;
    if (N < 1 || 4 * N > W) return cudaErrorInvalidValue;
    const float fr = box_radius(blur, passes);
    const int radius = (int)fr;
    const uint32_t ww = (uint32_t)((float)(1 << 24) / (fr * 2 + 1));
    const uint32_t fw = ((1u << 24) - (uint32_t)(radius * 2 + 1) * ww) / 2;
    cudaError_t e = cudaMemsetAsync(out, 0, (size_t)H * W * sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    const dim3 grid_c((unsigned)((size + 255) / 256), (unsigned)size), grid_f((unsigned)((N + 255) / 256), (unsigned)N);
    for (int f = 0; f < 6; ++f) {
        sp_blur_canvas_kernel<<<grid_c, 256, 0, st>>>(cross, tmp0, N, W, f);
        uint32_t *a = tmp0, *b = tmp1;
        for (int vertical = 0; vertical < 2; ++vertical)
            for (int p = 0; p < passes; ++p) {
                sp_blur_box_kernel<<<grid_c, 256, 0, st>>>(a, b, size, radius, ww, fw, vertical);
                uint32_t* t = a; a = b; b = t;
            }
        sp_blur_extract_kernel<<<grid_f, 256, 0, st>>>(a, out, N, W, f);
    }
    return cudaGetLastError();
}
