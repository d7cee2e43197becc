// GAE as a segmented affine REVERSE SCAN (BASELINE.json north_star: "warp shuffles handle the reverse-scan GAE over the
// time dimension"; src/ppo/data_loader.py:103-130).  Opt-in companion of g2048_gae_flat: that kernel walks every episode
// with the reference's exact fp32 operation order (bit-identical, 0.65-0.74 of HBM because a walking lane moves no
// bytes); this one re-associates the recurrence and is a pure stream -- results agree within the 1e-5 relative
// tolerance north_star states for GAE (tests/test_gae_scan_gpu.py), not bit for bit.
//
//   gae_t = delta_t + a_t * gae_{t+1},   a_t = done_t ? 0 : gamma * lambda,
//   delta_t = (r_t + gamma * (done_t ? 0 : V_{t+1})) - V_t
//
// i.e. gae_t = f_t(gae_{t+1}) with the affine map f_t(x) = a_t x + delta_t; maps compose associatively
// ((a1,b1) o (a2,b2) = (a1 a2, a1 b2 + b1)), so the suffix compositions F_t = f_t o f_{t+1} o ... are a scan:
//   * a thread composes its 8 consecutive steps serially (registers, 128-bit loads),
//   * a warp scans the 32 thread aggregates with shuffles (5 steps), the 8 warp aggregates go through shared memory,
//   * tiles (2 048 steps, taken from the END of the buffer by ticket by persistent CTAs) are chained by decoupled look-back:
//     a tile publishes its aggregate (A, B) and, once it knows the gae entering it, the gae of its first step.  A tile
//     that contains a `done` has A == 0 exactly, so its first-step gae is B and is published at once -- with episodes of
//     a few hundred steps the chain is never longer than one tile and no CTA waits for another's look-back.
// 9 B read + 8 B written per step; fp64 moment sums for the normalisation in the same pass.
#include "g2048_common.cuh"

namespace g2048 {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 2 048 steps
constexpr int SCAN_WARPS = SCAN_THREADS / 32;

struct ScanHeader {
    unsigned int ticket;
    unsigned int pad[3];
};
// per tile: word = flag << 32 | float bits (flag 1: B of the tile aggregate, A in `a`; flag 2: gae of the tile's first step)
struct ScanTile {
    unsigned long long word;
    float a;
    float pad;
};
constexpr unsigned SCAN_AGG = 1u, SCAN_INCL = 2u;

struct Affine {
    float a, b;
};
// apply `inner` first, then `outer`
__device__ __forceinline__ Affine compose(Affine outer, Affine inner) {
    return Affine{outer.a * inner.a, outer.a * inner.b + outer.b};
}

template <bool ALIGNED>
__global__ void __launch_bounds__(SCAN_THREADS, 4)
gae_scan_kernel(const float* __restrict__ rewards, const float* __restrict__ values, const uint8_t* __restrict__ dones,
                int64_t n, int64_t n_tiles, float gamma, float gamma_lambda, float* __restrict__ adv,
                float* __restrict__ ret, ScanHeader* header, double* __restrict__ moments) {
    __shared__ Affine s_warp[2][SCAN_WARPS];  // double-buffered by tile parity: one barrier per tile
    __shared__ unsigned int s_ticket[2];
    __shared__ double s_red[4 * SCAN_WARPS];

    ScanTile* tiles = reinterpret_cast<ScanTile*>(header + 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // Persistent CTAs: tiles are taken by ticket from the END of the buffer (a tile only ever waits for tiles with
    // earlier tickets); the ticket of the NEXT tile is drawn while the current one is being processed, so nobody waits
    // for that atomic (one CTA per tile did: `barrier` was the top stall, 14.7 per issue, 0.41 of HBM), and the moment
    // sums stay in registers until the CTA is done (5 atomics per CTA instead of per tile).
    if (threadIdx.x == 0) s_ticket[0] = atomicAdd(&header->ticket, 1u);
    __syncthreads();
    double acc4[4] = {0.0, 0.0, 0.0, 0.0};
    for (unsigned it = 0;; ++it) {
        const unsigned ticket = s_ticket[it & 1u];
        if ((int64_t)ticket >= n_tiles) break;  // uniform
        unsigned next_ticket = 0u;
        if (threadIdx.x == 0) next_ticket = atomicAdd(&header->ticket, 1u);  // in flight while this tile is loaded
        const int64_t tile = n_tiles - 1 - (int64_t)ticket;  // memory-order index: later tiles start first
        const int64_t lo = tile * SCAN_TILE;
        const int64_t first = lo + (int64_t)threadIdx.x * SCAN_ITEMS;  // this thread's first step

        // ---- loads: 8 consecutive steps per thread ---------------------------------------------------------------
        float r[SCAN_ITEMS], v[SCAN_ITEMS + 1];
        bool d[SCAN_ITEMS];
        const bool full = first + SCAN_ITEMS <= n;
        if (ALIGNED && full) {
            const float4 r0 = *reinterpret_cast<const float4*>(rewards + first), r1 = *reinterpret_cast<const float4*>(rewards + first + 4);
            const float4 v0 = *reinterpret_cast<const float4*>(values + first), v1 = *reinterpret_cast<const float4*>(values + first + 4);
            const uint2 dd = *reinterpret_cast<const uint2*>(dones + first);
            r[0] = r0.x; r[1] = r0.y; r[2] = r0.z; r[3] = r0.w; r[4] = r1.x; r[5] = r1.y; r[6] = r1.z; r[7] = r1.w;
            v[0] = v0.x; v[1] = v0.y; v[2] = v0.z; v[3] = v0.w; v[4] = v1.x; v[5] = v1.y; v[6] = v1.z; v[7] = v1.w;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                d[k] = ((dd.x >> (8 * k)) & 0xFFu) != 0u;
                d[4 + k] = ((dd.y >> (8 * k)) & 0xFFu) != 0u;
            }
        } else {
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) {
                const bool in = first + k < n;
                r[k] = in ? rewards[first + k] : 0.0f;
                v[k] = in ? values[first + k] : 0.0f;
                d[k] = in ? dones[first + k] != 0 : true;  // past the end: a = 0, delta = 0 -- contributes nothing
            }
        }
        v[SCAN_ITEMS] = (first + SCAN_ITEMS < n) ? values[first + SCAN_ITEMS] : 0.0f;  // V of the step after mine (0 past the end)

        // ---- per-step maps and the thread's aggregate (composition of its 8 steps, first step outermost) ---------------
        float a[SCAN_ITEMS], b[SCAN_ITEMS];
        Affine mine{1.0f, 0.0f};
#pragma unroll
        for (int k = SCAN_ITEMS - 1; k >= 0; --k) {
            const float next_v = d[k] ? 0.0f : v[k + 1];
            a[k] = d[k] ? 0.0f : gamma_lambda;
            b[k] = (r[k] + gamma * next_v) - v[k];
            mine = compose(Affine{a[k], b[k]}, mine);
        }
        // ---- warp: inclusive suffix scan over lanes (lane L: composition of lanes L .. 31) ----------------------------
        Affine incl = mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            Affine other;
            other.a = __shfl_down_sync(0xFFFFFFFFu, incl.a, off);
            other.b = __shfl_down_sync(0xFFFFFFFFu, incl.b, off);
            if (lane + off < 32) incl = compose(incl, other);
        }
        if (lane == 0) s_warp[it & 1u][warp] = incl;
        if (threadIdx.x == 0) s_ticket[(it + 1u) & 1u] = next_ticket;
        // what follows my steps inside the warp: the inclusive value of the next lane
        Affine after;
        after.a = __shfl_down_sync(0xFFFFFFFFu, incl.a, 1);
        after.b = __shfl_down_sync(0xFFFFFFFFu, incl.b, 1);
        if (lane == 31) after = Affine{1.0f, 0.0f};
        __syncthreads();  // the one barrier of a tile: warp aggregates and the next ticket are visible after it
        // ... and the warps after mine
        Affine later{1.0f, 0.0f};
        for (int w = SCAN_WARPS - 1; w > warp; --w) later = compose(s_warp[it & 1u][w], later);
        after = compose(after, later);  // everything between my last step and the end of the tile

        // ---- chain the tiles: publish, look back -------------------------------------------------------------------
        // gae of the first step of the next tile (0 past the end of the buffer): the first-step gae of the first following
        // tile that knows it, pushed through the aggregates of the tiles in between.  Any thread may ask; tiles with a
        // later index hold earlier tickets, so they have been started and the wait is short.
        auto look_back = [&]() -> float {
            Affine acc{1.0f, 0.0f};
            for (int64_t j = tile + 1;; ++j) {
                if (j == n_tiles) return acc.b;  // ran off the end of the buffer: the gae entering it is 0
                volatile unsigned long long* w = &tiles[j].word;
                unsigned long long word;
                do {
                    word = *w;
                } while ((unsigned)(word >> 32) == 0u);
                const float val = __uint_as_float((unsigned)word);
                if ((unsigned)(word >> 32) == SCAN_INCL) return acc.a * val + acc.b;
                __threadfence();
                acc = compose(acc, Affine{*(volatile float*)&tiles[j].a, val});
            }
        };
        if (threadIdx.x == 0) {
            const Affine whole = compose(incl, later);  // thread 0: lanes 0..31 of warp 0, then the later warps
            volatile unsigned long long* my_word = &tiles[tile].word;
            if (whole.a == 0.0f || tile == n_tiles - 1) {
                // nothing of what follows reaches my first step (or nothing follows): its gae is known now
                *my_word = ((unsigned long long)SCAN_INCL << 32) | (unsigned long long)__float_as_uint(whole.b);
            } else {
                tiles[tile].a = whole.a;
                __threadfence();
                *my_word = ((unsigned long long)SCAN_AGG << 32) | (unsigned long long)__float_as_uint(whole.b);
                const float x0 = look_back();
                *my_word = ((unsigned long long)SCAN_INCL << 32) | (unsigned long long)__float_as_uint(whole.a * x0 + whole.b);
            }
        }
        // Only the steps after the tile's last `done` see the next tile at all (after.a != 0): with episodes of a few
        // hundred steps that is the last warp or two of the CTA; every other thread goes straight on to its stores.
        const float x = (after.a != 0.0f) ? look_back() : 0.0f;

        // ---- apply: the gae entering my steps, then the reference's own recurrence over them ------------------------
        float g = after.a * x + after.b;
        float o_adv[SCAN_ITEMS], o_ret[SCAN_ITEMS];
#pragma unroll
        for (int k = SCAN_ITEMS - 1; k >= 0; --k) {
            g = b[k] + a[k] * g;
            o_adv[k] = g;
            o_ret[k] = g + v[k];
            if (first + k < n) {
                acc4[0] += (double)g;
                acc4[1] += (double)g * (double)g;
                acc4[2] += (double)o_ret[k];
                acc4[3] += (double)o_ret[k] * (double)o_ret[k];
            }
        }
        if (ALIGNED && full) {
            *reinterpret_cast<float4*>(adv + first) = make_float4(o_adv[0], o_adv[1], o_adv[2], o_adv[3]);
            *reinterpret_cast<float4*>(adv + first + 4) = make_float4(o_adv[4], o_adv[5], o_adv[6], o_adv[7]);
            *reinterpret_cast<float4*>(ret + first) = make_float4(o_ret[0], o_ret[1], o_ret[2], o_ret[3]);
            *reinterpret_cast<float4*>(ret + first + 4) = make_float4(o_ret[4], o_ret[5], o_ret[6], o_ret[7]);
        } else {
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) {
                if (first + k < n) {
                    adv[first + k] = o_adv[k];
                    ret[first + k] = o_ret[k];
                }
            }
        }
    }
    if (moments) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double sum = acc4[k];
            for (int off = 16; off > 0; off >>= 1) sum += __shfl_down_sync(0xFFFFFFFFu, sum, off);
            if (lane == 0) s_red[k * SCAN_WARPS + warp] = sum;
        }
        __syncthreads();
        if (threadIdx.x < 4) {
            double sum = 0.0;
            for (int w = 0; w < SCAN_WARPS; ++w) sum += s_red[threadIdx.x * SCAN_WARPS + w];
            atomicAdd(&moments[1 + threadIdx.x], sum);
        }
        if (threadIdx.x == 4 && blockIdx.x == 0) atomicAdd(&moments[0], (double)n);
    }
}

}  // namespace g2048

using namespace g2048;

extern "C" int64_t g2048_gae_scan_scratch_bytes(int64_t n) {
    const int64_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    return (int64_t)sizeof(ScanHeader) + n_tiles * (int64_t)sizeof(ScanTile);
}

extern "C" int g2048_gae_flat_scan(const float* d_rewards, const float* d_values, const uint8_t* d_dones, int64_t n, double gamma,
                                   double lambda_gae, float* d_adv, float* d_ret, void* d_scan_state, double* d_moments,
                                   void* stream) {
    G2048_REQUIRE(n >= 0, "gae_flat_scan: n");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_rewards && d_values && d_dones && d_adv && d_ret && d_scan_state, "gae_flat_scan: pointers");
    G2048_REQUIRE(((uintptr_t)d_scan_state & 15u) == 0, "gae_flat_scan: scratch must be 16-byte aligned");
    const int64_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    const auto a16 = [](const void* p) { return ((uintptr_t)p & 15u) == 0; };
    const bool aligned = a16(d_rewards) && a16(d_values) && a16(d_adv) && a16(d_ret) && ((uintptr_t)d_dones & 7u) == 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = sm_count();
    if (sms <= 0) return fail_arg("gae_flat_scan: no device");
    const int64_t resident = (int64_t)sms * 4;  // persistent: four 256-thread CTAs per SM
    const unsigned grid = (unsigned)(n_tiles < resident ? n_tiles : resident);
    if (aligned)
        gae_scan_kernel<true><<<grid, SCAN_THREADS, 0, st>>>(d_rewards, d_values, d_dones, n, n_tiles, (float)gamma,
                                                                          (float)(gamma * lambda_gae), d_adv, d_ret,
                                                                          (ScanHeader*)d_scan_state, d_moments);
    else
        gae_scan_kernel<false><<<grid, SCAN_THREADS, 0, st>>>(d_rewards, d_values, d_dones, n, n_tiles, (float)gamma,
                                                                           (float)(gamma * lambda_gae), d_adv, d_ret,
                                                                           (ScanHeader*)d_scan_state, d_moments);
    G2048_CHECK_LAUNCH("gae_flat_scan");
    return G2048_OK;
}
