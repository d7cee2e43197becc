// Pieces shared by the persistent play kernels (g2048_play.cu: SWAR board logic, g2048_play3.cu: row tables).
#pragma once
#include "g2048_board.cuh"
#include "g2048_env.cuh"

namespace g2048 {

// The blocks of one key that both bits4() and split2() are made of.
template <int MODE>
struct KeyBlocks {
    uint32_t bits[4];  // random_bits(key, (4,))
    Key child[2];      // split(key, 2)
};

template <int MODE>
__device__ __forceinline__ KeyBlocks<MODE> key_blocks(Key k) {
    KeyBlocks<MODE> o;
    if (MODE == G2048_RNG_PARTITIONABLE) {
        Key y[4];
#pragma unroll
        for (uint32_t i = 0; i < 4; ++i) {
            y[i] = threefry2x32(k, 0u, i);
            o.bits[i] = y[i].a ^ y[i].b;
        }
        o.child[0] = y[0];
        o.child[1] = y[1];
    } else {
        const Key y0 = threefry2x32(k, 0u, 2u);
        const Key y1 = threefry2x32(k, 1u, 3u);
        o.bits[0] = y0.a;
        o.bits[1] = y1.a;
        o.bits[2] = y0.b;
        o.bits[3] = y1.b;
        o.child[0] = Key{y0.a, y1.a};
        o.child[1] = Key{y0.b, y1.b};
    }
    return o;
}

__device__ __forceinline__ int argmax_bits_legal(const uint32_t bits[4], uint32_t legal) {
    const uint32_t allowed = (legal & 15u) ? (legal & 15u) : 15u;
    int best = 0, best_v = -1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int v = ((allowed >> i) & 1u) ? (int)(bits[i] >> 9) : -1;
        if (v > best_v) {
            best_v = v;
            best = i;
        }
    }
    return best;
}

enum : uint32_t { PHASE_INIT0 = 0, PHASE_INIT1 = 1, PHASE_PLAY = 2, PHASE_NONE = 3 };

}  // namespace g2048
