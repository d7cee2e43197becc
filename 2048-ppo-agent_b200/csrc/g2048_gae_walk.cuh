// Device helpers shared by the flat-GAE kernels (g2048_gae3.cu, g2048_gae4.cu): the in-place backward walk over one
// episode's deltas in shared memory, the ordered-done-list lookup, and the 128-bit loads of phase 1.
#pragma once
#include <cstdint>

namespace g2048 {

// bits 0..7 of b -> bit positions 0, 4, 8, ..., 28
__device__ __forceinline__ uint32_t spread_bits4(uint32_t b) {
    uint32_t x = b & 0xFFu;
    x = (x | (x << 12)) & 0x000F000Fu;
    x = (x | (x << 6)) & 0x03030303u;
    x = (x | (x << 3)) & 0x11111111u;
    return x;
}

// position (in the tile) of the k-th done step, k < n_done.  ballot: four words per 128-step block (ballot j has bit
// l set iff step 128*b + 4*l + j is a done); pref: exclusive prefix of the blocks' done counts.  Runs once per
// episode, so the bit interleave of the four ballots is done here and not by every warp in phase 1.
__device__ __forceinline__ int gae_locate(const uint32_t* __restrict__ ballot, const uint32_t* __restrict__ pref,
                                          int n_blocks, int k) {
    int lo = 0, hi = n_blocks - 1;  // largest b with pref[b] <= k
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((int)pref[mid] <= k) lo = mid; else hi = mid - 1;
    }
    const uint32_t b0 = ballot[4 * lo], b1 = ballot[4 * lo + 1], b2 = ballot[4 * lo + 2], b3 = ballot[4 * lo + 3];
    int r = k - (int)pref[lo];
    for (int c = 0; c < 4; ++c) {  // 32 steps (8 lanes x 4 components) at a time
        const int sh = 8 * c;
        uint32_t m = spread_bits4(b0 >> sh) | (spread_bits4(b1 >> sh) << 1) | (spread_bits4(b2 >> sh) << 2) |
                     (spread_bits4(b3 >> sh) << 3);
        const int cnt = __popc(m);
        if (r < cnt) {
            for (; r > 0; --r) m &= m - 1;  // drop the r lowest set bits
            return 128 * lo + 32 * c + (__ffs((int)m) - 1);
        }
        r -= cnt;
    }
    return 128 * n_blocks - 1;  // not reached for k < n_done
}

// gae = delta + gl * gae backwards over the steps (first_excl, last], in place
__device__ __forceinline__ void gae_walk(float* __restrict__ sg, int last, int first_excl, float g, float gl_in) {
    // gamma*lambda in a register of its own: left alone, ptxas re-loads the kernel parameter from the constant bank
    // at the top of every trip (LDC) and the first multiply of the serial chain waits for it
    float gl;
    asm volatile("mov.f32 %0, %1;" : "=f"(gl) : "f"(gl_in));
    int t = last;
    while (t > first_excl && (t & 3) != 3) {  // down to a 16-byte boundary
        g = sg[t] + gl * g;
        sg[t] = g;
        --t;
    }
    // 4-step groups, 16 steps per trip.  ncu on the previous form (two 8-step half trips with swapped register
    // sets) showed ptxas merging the halves back into one 8-step body with 12 register moves and 5.5 instructions
    // per step; here every group of the trip has its own offset and its own registers, so there is nothing to
    // merge: 4 LDS.128 + 16 FMUL + 16 FADD + 4 STS.128 + loop control.  The loads of the second half are issued
    // before the first half's chain, those of the next trip's first half before the second half's chain; an
    // output set is rewritten two groups after its store was issued (STS.128 holds its sources until dispatched).
#define GAE_LD(at) (*reinterpret_cast<const float4*>(&sg[(at)]))
#define GAE_GROUP(in, out, at)                                  \
    g = in.w + gl * g; out.w = g;                                \
    g = in.z + gl * g; out.z = g;                                \
    g = in.y + gl * g; out.y = g;                                \
    g = in.x + gl * g; out.x = g;                                \
    *reinterpret_cast<float4*>(&sg[(at)]) = out;
    if (t - 16 >= first_excl) {  // steps t-15 .. t are all inside the episode
        float4 a0 = GAE_LD(t - 3), a1 = GAE_LD(t - 7), b0, b1, o0, o1;
#pragma unroll 1
        do {
            b0 = GAE_LD(t - 11);
            b1 = GAE_LD(t - 15);
            GAE_GROUP(a0, o0, t - 3)
            GAE_GROUP(a1, o1, t - 7)
            const bool more = t - 32 >= first_excl;  // another full trip follows: fetch its first half now
            if (more) {
                a0 = GAE_LD(t - 19);
                a1 = GAE_LD(t - 23);
            }
            GAE_GROUP(b0, o0, t - 11)
            GAE_GROUP(b1, o1, t - 15)
            t -= 16;
            if (!more) break;
        } while (true);
    }
    while (t - 4 >= first_excl) {  // at most three groups remain
        const float4 in = GAE_LD(t - 3);
        float4 out;
        GAE_GROUP(in, out, t - 3)
        t -= 4;
    }
#undef GAE_GROUP
#undef GAE_LD
    for (; t > first_excl; --t) {
        g = sg[t] + gl * g;
        sg[t] = g;
    }
}

template <bool ALIGNED>
__device__ __forceinline__ float4 gae_load4(const float* __restrict__ p, int64_t gi, int valid) {
    if (ALIGNED && valid == 4) return __ldg(reinterpret_cast<const float4*>(p + gi));
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid > 0) o.x = __ldg(p + gi);
    if (valid > 1) o.y = __ldg(p + gi + 1);
    if (valid > 2) o.z = __ldg(p + gi + 2);
    if (valid > 3) o.w = __ldg(p + gi + 3);
    return o;
}

// L2 eviction policies (createpolicy): V is read twice, once for delta and about one pipeline iteration later for
// ret = adv + V; ncu showed two thirds of the second reads going back to DRAM (11.6 B read per step instead of 9),
// so V is loaded "evict last" the first time and "evict first" the second, like the streams that are read once.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

template <bool ALIGNED>
__device__ __forceinline__ float4 gae_load4_hint(const float* __restrict__ p, int64_t gi, int valid, uint64_t policy) {
    if (ALIGNED && valid == 4) {
        float4 o;
        asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                     : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                     : "l"(p + gi), "l"(policy));
        return o;
    }
    return gae_load4<ALIGNED>(p, gi, valid);
}

template <bool ALIGNED>
__device__ __forceinline__ uint32_t gae_load_done4(const uint8_t* __restrict__ p, int64_t gi, int valid) {
    uint32_t w = 0;
    if (ALIGNED && valid == 4) {
        w = __ldg(reinterpret_cast<const uint32_t*>(p + gi));
    } else {
        if (valid > 0) w |= (uint32_t)__ldg(p + gi);
        if (valid > 1) w |= (uint32_t)__ldg(p + gi + 1) << 8;
        if (valid > 2) w |= (uint32_t)__ldg(p + gi + 2) << 16;
        if (valid > 3) w |= (uint32_t)__ldg(p + gi + 3) << 24;
    }
    return w;
}

}  // namespace g2048
