// TEST-ONLY host definitions of the CUDA intrinsics the g2048 device headers use, so that the
// per-env device functions (bitboard move, spawn, legal mask, Threefry, policies) can be compiled
// with g++ and unit-tested against the oracle in the CPU-only container.  Nothing in the product
// (libg2048.so / the g2048 package) includes this file; GPU parity is proven by the -m gpu tests.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

using std::max;
using std::min;

struct uint2 { uint32_t x, y; };
struct float4 { float x, y, z, w; };

static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, int s) {
    s &= 31;
    return s ? (hi << s) | (lo >> (32 - s)) : hi;
}
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
    uint8_t src[8];
    std::memcpy(src, &a, 4);
    std::memcpy(src + 4, &b, 4);
    uint32_t out = 0;
    for (int i = 0; i < 4; ++i) out |= (uint32_t)src[(sel >> (4 * i)) & 7] << (8 * i);
    return out;
}
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
