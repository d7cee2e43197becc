// Device code shared by the flat-GAE kernels (g2048_gae3.cu, g2048_gae4.cu): the in-place backward walk over one
// episode's deltas in shared memory, the ordered-done-list lookup, the hinted 128-bit loads, and the three tile
// phases (deltas + done ballots, block prefix, store + moments).
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


// ------------------------------------------------------------------------------------------------------------------
// Tile phases shared by the one-tile-per-CTA kernel (g2048_gae3.cu) and the pipelined one (g2048_gae4.cu).
// THREADS threads cooperate on a TILE-step tile; thread t owns the 4 consecutive steps 4*(q*THREADS + t) .. +3 of
// group q, so a warp covers one 128-step block per group; INFLIGHT groups are loaded before the first is used.
// ------------------------------------------------------------------------------------------------------------------

// inputs of tile `pt` (a full tile) into L2, one warp; V with evict-last priority because it is read twice
template <int TILE>
__device__ __forceinline__ void gae_tile_prefetch(const float* __restrict__ rewards, const float* __restrict__ values,
                                                  const uint8_t* __restrict__ dones, int64_t pt, int lane) {
    const char* pr = reinterpret_cast<const char*>(rewards + pt * TILE);
    const char* pv = reinterpret_cast<const char*>(values + pt * TILE);
    const char* pd = reinterpret_cast<const char*>(dones + pt * TILE);
    for (int i = lane * 128; i < TILE * 4; i += 32 * 128) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pr + i));
        asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(pv + i));
    }
    for (int i = lane * 128; i < TILE; i += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pd + i));
}

// phase 1: delta = (r + gamma * V[t+1] * !done) - V into g[], the four done ballots of every 128-step block into
// ballot[] (ballot j of block b has bit l set iff step 128*b + 4*l + j is a done).  lo = first step of the tile,
// len = its length (< TILE only for the last tile of the buffer).
template <bool ALIGNED, int TILE, int THREADS, int INFLIGHT>
__device__ __forceinline__ void gae_tile_deltas(float* __restrict__ g, uint32_t* __restrict__ ballot,
                                                const float* __restrict__ rewards, const float* __restrict__ values,
                                                const uint8_t* __restrict__ dones, int64_t n, int64_t lo, int len, float gamma,
                                                int tid) {
    constexpr int VEC = TILE / (4 * THREADS), PASSES = VEC / INFLIGHT, WARPS = THREADS / 32;
    static_assert(VEC * 4 * THREADS == TILE && PASSES * INFLIGHT == VEC, "tile shape");
    const int lane = tid & 31, warp = tid >> 5;
    const uint64_t keep = l2_policy_evict_last(), once = l2_policy_evict_first();
#pragma unroll
    for (int h = 0; h < PASSES; ++h) {
        float4 r[INFLIGHT], v[INFLIGHT];
        uint32_t d[INFLIGHT];
        float vnext[INFLIGHT];
#pragma unroll
        for (int k = 0; k < INFLIGHT; ++k) {
            const int i = 4 * ((h * INFLIGHT + k) * THREADS + tid);
            const int valid = max(0, min(4, len - i));
            r[k] = gae_load4_hint<ALIGNED>(rewards, lo + i, valid, once);
            v[k] = gae_load4_hint<ALIGNED>(values, lo + i, valid, keep);  // read again when the tile is stored
            d[k] = gae_load_done4<ALIGNED>(dones, lo + i, valid);
            // V of the step after this lane's four: the next lane has it, except for lane 31
            vnext[k] = (lane == 31 && lo + i + 4 < n && i + 4 <= len + 3) ? __ldg(values + lo + i + 4) : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < INFLIGHT; ++k) {
            const int q = h * INFLIGHT + k;
            const int i = 4 * (q * THREADS + tid);
            const float from_next_lane = __shfl_down_sync(0xFFFFFFFFu, v[k].x, 1);
            const float v4 = (lane == 31) ? vnext[k] : from_next_lane;  // 0 past the end of the buffer
            const bool d0 = (d[k] & 0xFFu) != 0, d1 = (d[k] & 0xFF00u) != 0, d2 = (d[k] & 0xFF0000u) != 0,
                       d3 = (d[k] & 0xFF000000u) != 0;
            float4 delta;
            delta.x = (r[k].x + gamma * (d0 ? 0.0f : v[k].y)) - v[k].x;
            delta.y = (r[k].y + gamma * (d1 ? 0.0f : v[k].z)) - v[k].y;
            delta.z = (r[k].z + gamma * (d2 ? 0.0f : v[k].w)) - v[k].z;
            delta.w = (r[k].w + gamma * (d3 ? 0.0f : v4)) - v[k].w;
            *reinterpret_cast<float4*>(&g[i]) = delta;
            const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, d0), b1 = __ballot_sync(0xFFFFFFFFu, d1),
                           b2 = __ballot_sync(0xFFFFFFFFu, d2), b3 = __ballot_sync(0xFFFFFFFFu, d3);
            if (lane < 4)  // as they are: gae_locate interleaves them when an episode looks its end up
                ballot[4 * (q * WARPS + warp) + lane] = lane == 0 ? b0 : (lane == 1 ? b1 : (lane == 2 ? b2 : b3));
        }
    }
}

// exclusive prefix over the blocks' done counts, by ONE warp after all ballots are written; pref[BLOCKS] = n_done
template <int BLOCKS>
__device__ __forceinline__ void gae_tile_prefix(const uint32_t* __restrict__ ballot, uint32_t* __restrict__ pref, int lane) {
    constexpr int PER_LANE = (BLOCKS + 31) / 32;
    uint32_t c[PER_LANE];
    uint32_t sum = 0;
#pragma unroll
    for (int q = 0; q < PER_LANE; ++q) {
        const int blk = lane * PER_LANE + q;
        c[q] = blk < BLOCKS ? __popc(ballot[4 * blk]) + __popc(ballot[4 * blk + 1]) + __popc(ballot[4 * blk + 2]) +
                                  __popc(ballot[4 * blk + 3])
                            : 0u;
        sum += c[q];
    }
    uint32_t incl = sum;
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, off);
        if (lane >= off) incl += y;
    }
    uint32_t run = incl - sum;
#pragma unroll
    for (int q = 0; q < PER_LANE; ++q) {
        const int blk = lane * PER_LANE + q;
        if (blk < BLOCKS) pref[blk] = run;
        run += c[q];
    }
    if (lane == 31) pref[BLOCKS] = incl;
}

// phase 3: advantages from g[], V again from global (an L2 hit: loaded evict-last in phase 1, released here),
// returns = adv + V, 128-bit streaming stores, fp64 moment sums m = {sum adv, sum adv^2, sum ret, sum ret^2}
template <bool ALIGNED, int TILE, int THREADS, int INFLIGHT>
__device__ __forceinline__ void gae_tile_store(const float* __restrict__ g, const float* __restrict__ values, int64_t lo,
                                               int len, float* __restrict__ adv, float* __restrict__ ret, double (&m)[4],
                                               int tid) {
    constexpr int VEC = TILE / (4 * THREADS), PASSES = VEC / INFLIGHT;
    const uint64_t once = l2_policy_evict_first();
#pragma unroll
    for (int h = 0; h < PASSES; ++h) {
        float4 v[INFLIGHT];
#pragma unroll
        for (int k = 0; k < INFLIGHT; ++k) {
            const int i = 4 * ((h * INFLIGHT + k) * THREADS + tid);
            v[k] = gae_load4_hint<ALIGNED>(values, lo + i, max(0, min(4, len - i)), once);
        }
#pragma unroll
        for (int k = 0; k < INFLIGHT; ++k) {
            const int i = 4 * ((h * INFLIGHT + k) * THREADS + tid);
            const int valid = max(0, min(4, len - i));
            if (valid > 0) {
                const float4 a = *reinterpret_cast<const float4*>(&g[i]);
                const float4 rt = make_float4(a.x + v[k].x, a.y + v[k].y, a.z + v[k].z, a.w + v[k].w);
                const float aa[4] = {a.x, a.y, a.z, a.w}, rr[4] = {rt.x, rt.y, rt.z, rt.w};
                if (ALIGNED && valid == 4) {
                    __stcs(reinterpret_cast<float4*>(adv + lo + i), a);
                    __stcs(reinterpret_cast<float4*>(ret + lo + i), rt);
                } else {
                    for (int j = 0; j < valid; ++j) {
                        adv[lo + i + j] = aa[j];
                        ret[lo + i + j] = rr[j];
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (j < valid) {
#ifndef G2048_GAE_SKIP_MOMENTS  // timing experiment only (tools/probes)
                        const double da = (double)aa[j], dr = (double)rr[j];
                        m[0] += da;
                        m[1] = __fma_rn(da, da, m[1]);  // the product of two floats is exact in double either way
                        m[2] += dr;
                        m[3] = __fma_rn(dr, dr, m[3]);
#endif
                    }
                }
            }
        }
    }
}

}  // namespace g2048
