// Counter-based random numbers: Philox4x32-10 keyed by the frame seed, countered by
// (pixel, path-tree node, draw block).  Mirrored bit for bit by oracle/sightpy_oracle.py
// (philox4x32 / child_path / root_path / u01), which is what makes Monte-Carlo scenes comparable
// ray by ray.  Replaces the numpy global stream of the reference (camera.py:56-61, random.py).
#pragma once
#include "sp_math.cuh"

#define SP_BLOCK_DIRECTION 0u   // draws that generate this node's ray (camera jitter / diffuse direction)
#define SP_BLOCK_MATERIAL  1u   // draws made while shading this node's hit (Refractive mc pick)

SP_DEV void sp_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 24-bit uniform in [0,1): exactly representable in float32 (and identical in the float64 oracle)
SP_DEV float sp_u01(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }

// hash of the k-th child of a node of the path tree
SP_DEV uint32_t sp_child_path(uint32_t path, uint32_t k) {
    uint32_t x = (path * 0x01000193u) ^ ((k + 1u) * 0x9E3779B9u);
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
}
SP_DEV uint32_t sp_root_path(uint32_t sample) { return sp_child_path(0x811C9DC5u, sample); }

// The same generator with the ten round keys (k0 + r * 0x9E3779B9, k1 + r * 0xBB67AE85) read from a table the
// host fills per call (DScene::philox_keys, kernel parameter space): the key schedule costs no instructions.
SP_DEV void sp_philox4x32_10_keys(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t* __restrict__ keys,
                                   uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ keys[2 * r]; c1 = lo1; c2 = hi0 ^ c3 ^ keys[2 * r + 1]; c3 = lo0;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

SP_DEV void sp_draw4_keys(uint32_t pix, uint32_t path, uint32_t block, const uint32_t* __restrict__ keys, float u[4]) {
    uint32_t w[4];
    sp_philox4x32_10_keys(pix, path, block, 0u, keys, w);
    u[0] = sp_u01(w[0]); u[1] = sp_u01(w[1]); u[2] = sp_u01(w[2]); u[3] = sp_u01(w[3]);
}

SP_DEV void sp_draw4(uint32_t pix, uint32_t path, uint32_t block, uint32_t k0, uint32_t k1, float u[4]) {
    uint32_t w[4];
    sp_philox4x32_10(pix, path, block, 0u, k0, k1, w);
    u[0] = sp_u01(w[0]); u[1] = sp_u01(w[1]); u[2] = sp_u01(w[2]); u[3] = sp_u01(w[3]);
}
