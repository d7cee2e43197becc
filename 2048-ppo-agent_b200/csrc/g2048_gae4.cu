// GAE over a flat buffer, pipelined generation (src/ppo/data_loader.py:103-130; bit-identical results).
//
// Experiments on g2048_gae3.cu (tools/probes/probe_gae3.cu, 2^26 steps, B200): with every walk removed the kernel
// still needs 244 us (4.7 TB/s) and the walks add 70 us on top, because inside a CTA the three phases run one after
// the other: while a tile's longest episode is walked the CTA moves no bytes, and while it loads or stores nobody
// walks.  Here a CTA is persistent and keeps TWO tiles in shared memory:
//
//     iteration i :  walker warps   walk tile i          (stage i % 2)
//                    streamer warps store tile i-1, then load tile i+1 into the freed stage ((i+1) % 2)
//
// so the serial recurrence of tile i is hidden behind the HBM traffic of its neighbours.  Tiles are still taken by
// ticket from the end of the buffer, a tile's open tail still waits for the head of the tile after it (decoupled
// look-back of depth one), and every CTA walks its tiles in ticket order, so the earliest unfinished ticket never
// waits on a later one.  Eight streamer warps (256 threads, the phase-1 / phase-3 code of the previous generation)
// + four walker warps (two for whole episodes, one for the tile's first episode, one for its tail).
#include <cstdlib>

#include "g2048_common.cuh"
#include "g2048_gae_walk.cuh"

namespace g2048 {

#ifndef G4_STREAMERS
#define G4_STREAMERS 256
#endif
constexpr int G4_STREAM_THREADS = G4_STREAMERS;
constexpr int G4_STREAM_WARPS = G4_STREAM_THREADS / 32;
#ifndef G4_WALKER_WARPS
#define G4_WALKER_WARPS 4
#endif
constexpr int G4_WALK_WARPS = G4_WALKER_WARPS;  // 4: two for whole episodes, first episode, tail; 2: whole episodes, first episode then tail
static_assert(G4_WALK_WARPS == 4 || G4_WALK_WARPS == 2, "two or four walker warps");
constexpr int G4_THREADS = G4_STREAM_THREADS + 32 * G4_WALK_WARPS;  // 384
#ifndef G4_TILE_STEPS
#define G4_TILE_STEPS 8192
#endif
#ifndef G4_MIN_CTAS
#define G4_MIN_CTAS 3
#endif
constexpr int G4_TILE = G4_TILE_STEPS;
constexpr int G4_BLOCKS = G4_TILE / 128;
#ifndef G4_STORE_GROUPS_INFLIGHT
#define G4_STORE_GROUPS_INFLIGHT 2
#endif
constexpr int G4_STORE_INFLIGHT = G4_STORE_GROUPS_INFLIGHT;  // the same for the store phase, which only loads V
#ifndef G4_GROUPS_INFLIGHT
#define G4_GROUPS_INFLIGHT 2
#endif
constexpr int G4_INFLIGHT = G4_GROUPS_INFLIGHT;  // float4 groups a streamer thread loads before it uses the first (4: 56 registers spill)

struct G4Scratch {  // same layout as the previous generation: ticket, then flags[n_tiles], heads[n_tiles]
    unsigned int ticket;
    unsigned int pad[3];
};

struct G4Stage {
    float g[G4_TILE];                // delta -> advantages, in place
    uint32_t ballot[4 * G4_BLOCKS];  // done bits, four ballots per 128-step block
    uint32_t pref[G4_BLOCKS + 1];    // exclusive prefix of the blocks' done counts
    uint32_t pad[3 - (G4_BLOCKS % 4)];
};
static_assert(sizeof(G4Stage) % 16 == 0, "the second stage's deltas must stay 16-byte aligned");

struct G4Smem {
    G4Stage st[2];
    double red[4 * G4_STREAM_WARPS];
    // tile_ring[i & 3] = tile walked in iteration i (-1: none), written during iteration i - 1.  Four slots: a thread
    // still reading the slots of iterations i and i - 1 never meets the write for iteration i + 1.
    int tile_ring[4];
    int next_tile;
};

__device__ __forceinline__ void g4_stream_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(G4_STREAM_THREADS) : "memory"); }

// phase 1 of one tile (streamer threads only): deltas into st.g, done ballots, block prefix
template <bool ALIGNED>
__device__ __forceinline__ void g4_load_tile(G4Stage& st, int64_t tile, const float* __restrict__ rewards,
                                             const float* __restrict__ values, const uint8_t* __restrict__ dones, int64_t n,
                                             float gamma, int tid) {
    const int64_t lo = tile * G4_TILE;
    const int len = (int)min((int64_t)G4_TILE, n - lo);
    gae_tile_deltas<ALIGNED, G4_TILE, G4_STREAM_THREADS, G4_INFLIGHT>(st.g, st.ballot, rewards, values, dones, n, lo, len, gamma, tid);
    g4_stream_barrier();
    if (tid < 32) gae_tile_prefix<G4_BLOCKS>(st.ballot, st.pref, tid);
}

// phase 3 of one tile (streamer threads only), right at the start of the iteration after the walk: measured, lines
// survive about one iteration in L2 -- a variant that spread V's second read over the whole iteration, fused with
// the next tile's loads, was 13 % slower
template <bool ALIGNED>
__device__ __forceinline__ void g4_store_tile(const G4Stage& st, int64_t tile, const float* __restrict__ values, int64_t n,
                                              float* __restrict__ adv, float* __restrict__ ret, double (&m)[4], int tid) {
    const int64_t lo = tile * G4_TILE;
    const int len = (int)min((int64_t)G4_TILE, n - lo);
    gae_tile_store<ALIGNED, G4_TILE, G4_STREAM_THREADS, G4_STORE_INFLIGHT>(st.g, values, lo, len, adv, ret, m, tid);
}

// phase 2 of one tile (walker warps only; wwarp = 0..3)
__device__ __forceinline__ void g4_walk_tile(G4Stage& st, int64_t tile, int64_t n, float gamma_lambda,
                                             volatile unsigned int* flags, volatile float* heads, int wwarp, int lane) {
    const int64_t lo = tile * G4_TILE;
    const int len = (int)min((int64_t)G4_TILE, n - lo);
    const int n_done = (int)st.pref[G4_BLOCKS];
    const bool does_first = wwarp == 1, does_tail = G4_WALK_WARPS == 4 ? wwarp == 2 : wwarp == 1;
    if (does_first || does_tail) {
        if (lane == 0) {
            // the first episode of the tile: the previous tile is waiting for its result
            if (does_first && n_done > 0) {
#ifndef G4_SKIP_WALKS  // timing experiment only (tools/probes/probe_gae4.cu)
                gae_walk(st.g, gae_locate(st.ballot, st.pref, G4_BLOCKS, 0), -1, 0.0f, gamma_lambda);
#endif
                race_jitter();
                heads[tile] = st.g[0];
                __threadfence();
                race_jitter();
                flags[tile] = 1u;
            }
            // the steps after the tile's last done belong to an episode that ends in a later tile
            if (does_tail) {
                const int first_excl = n_done ? gae_locate(st.ballot, st.pref, G4_BLOCKS, n_done - 1) : -1;
                if (first_excl < len - 1) {
                    float carry = 0.0f;
                    if (lo + len < n) {
                        race_jitter();
                        while (flags[tile + 1] == 0u) __nanosleep(100);
                        __threadfence();
                        race_jitter();
                        carry = heads[tile + 1];
                    }
#ifndef G4_SKIP_WALKS
                    gae_walk(st.g, len - 1, first_excl, carry, gamma_lambda);
#else
                    (void)carry;
#endif
                }
                if (n_done == 0) {
                    race_jitter();
                    heads[tile] = st.g[0];
                    __threadfence();
                    race_jitter();
                    flags[tile] = 1u;
                }
            }
        }
    } else {
        constexpr int slots = G4_WALK_WARPS == 4 ? 2 : 1;
        const int slot = wwarp == 0 ? 0 : 1;
        for (int base = 1 + 32 * slot; base < n_done; base += slots * 32) {  // warp-uniform
            const int e = base + lane;
            const int end = e < n_done ? gae_locate(st.ballot, st.pref, G4_BLOCKS, e) : 0;
            int prev = __shfl_up_sync(0xFFFFFFFFu, end, 1);  // the episode before mine ends where my neighbour's does
            if (lane == 0) prev = gae_locate(st.ballot, st.pref, G4_BLOCKS, base - 1);
#ifndef G4_SKIP_WALKS
            if (e < n_done) gae_walk(st.g, end, prev, 0.0f, gamma_lambda);
#endif
        }
    }
}

template <bool ALIGNED>
__global__ void __launch_bounds__(G4_THREADS, G4_MIN_CTAS)
gae_flat4_kernel(const float* __restrict__ rewards, const float* __restrict__ values, const uint8_t* __restrict__ dones,
                 int64_t n, int64_t n_tiles, float gamma, float gamma_lambda, float* __restrict__ adv,
                 float* __restrict__ ret, G4Scratch* scratch, double* __restrict__ moments, int prefetch_tiles) {
    extern __shared__ __align__(16) unsigned char g4_smem_raw[];
    G4Smem& s = *reinterpret_cast<G4Smem*>(g4_smem_raw);
    volatile unsigned int* flags = (volatile unsigned int*)(scratch + 1);
    volatile float* heads = (volatile float*)((unsigned int*)(scratch + 1) + n_tiles);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool streamer = warp < G4_STREAM_WARPS;

    auto take_ticket = [&]() {  // tiles are taken from the END of the buffer; -1 when none is left
        race_jitter();
        const unsigned int t = atomicAdd(&scratch->ticket, 1u);
        s.next_tile = (int64_t)t < n_tiles ? (int)(n_tiles - 1 - (int64_t)t) : -1;
    };
    auto prefetch = [&](int64_t tile) {  // inputs of the tile `prefetch_tiles` tickets ahead into L2 (one warp)
        const int64_t pt = tile - prefetch_tiles;
        if (prefetch_tiles > 0 && pt >= 0) gae_tile_prefetch<G4_TILE>(rewards, values, dones, pt, lane);
    };

    // ---- prologue: the first tile into stage 0 ---------------------------------------------------------------
    if (tid == 0) {
        take_ticket();
        s.tile_ring[3] = -1;  // "iteration -1" walked nothing
    }
    __syncthreads();
    if (streamer) {
        const int first = s.next_tile;
        if (first >= 0) {
            if (warp == G4_STREAM_WARPS - 1) prefetch(first);
            g4_load_tile<ALIGNED>(s.st[0], first, rewards, values, dones, n, gamma, tid);
        }
        if (tid == 0) s.tile_ring[0] = first;
    }
    __syncthreads();

    double m[4] = {0.0, 0.0, 0.0, 0.0};
    for (int it = 0;; ++it) {
        G4Stage& cur = s.st[it & 1];
        G4Stage& other = s.st[(it & 1) ^ 1];
        const int tile_cur = s.tile_ring[it & 3];         // walked in this iteration
        const int tile_prev = s.tile_ring[(it + 3) & 3];  // walked in the previous one: stored now, its stage refilled
        if (tile_cur < 0 && tile_prev < 0) break;
        if (streamer) {
            if (tid == 0) {  // the ticket's round trip hides behind the stores
                if (tile_cur >= 0) take_ticket(); else s.next_tile = -1;
            }
            if (tile_prev >= 0) g4_store_tile<ALIGNED>(other, tile_prev, values, n, adv, ret, m, tid);
            g4_stream_barrier();  // next_tile is visible
            const int next = s.next_tile;
            if (next >= 0) {
                if (warp == G4_STREAM_WARPS - 1) prefetch(next);
                g4_load_tile<ALIGNED>(other, next, rewards, values, dones, n, gamma, tid);
            }
            if (tid == 0) s.tile_ring[(it + 1) & 3] = next;
        } else if (tile_cur >= 0) {
            g4_walk_tile(cur, tile_cur, n, gamma_lambda, flags, heads, warp - G4_STREAM_WARPS, lane);
        }
        __syncthreads();
    }

    if (moments && streamer) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double x = m[k];
            for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xFFFFFFFFu, x, off);
            if (lane == 0) s.red[k * G4_STREAM_WARPS + warp] = x;
        }
        g4_stream_barrier();
        if (tid < 4) {
            double t = 0.0;
            for (int w = 0; w < G4_STREAM_WARPS; ++w) t += s.red[tid * G4_STREAM_WARPS + w];
            atomicAdd(&moments[1 + tid], t);
        }
    }
}

// moments[0] = n: one thread, so that the count does not depend on how tiles were distributed
__global__ void gae_flat4_count_kernel(double* moments, double n) { atomicAdd(&moments[0], n); }

}  // namespace g2048

using namespace g2048;

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

extern "C" int g2048_gae_flat_pipelined(const float* d_rewards, const float* d_values, const uint8_t* d_dones, int64_t n,
                                        double gamma, double lambda_gae, float* d_adv, float* d_ret, void* d_scan_state,
                                        double* d_moments, void* stream) {
    G2048_REQUIRE(n >= 0, "gae_flat: n");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_rewards && d_values && d_dones && d_adv && d_ret && d_scan_state, "gae_flat: pointers");
    const int64_t n_tiles = (n + G4_TILE - 1) / G4_TILE;
    G2048_REQUIRE(n_tiles <= 0x7FFFFFFF, "gae_flat: too many tiles");
    const bool aligned = aligned16(d_rewards) && aligned16(d_values) && aligned16(d_adv) && aligned16(d_ret) &&
                         ((uintptr_t)d_dones & 3u) == 0;
    const int sms = sm_count();
    if (sms <= 0) return fail_arg("gae_flat: no device");
    static bool configured_on[64] = {false};
    bool* configured = device_once_flag(configured_on);
    if (!configured) return fail_arg("no CUDA device");
    if (!*configured) {
        int rc = check_cuda(cudaFuncSetAttribute(gae_flat4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(G4Smem)), "gae_flat: smem attribute");
        if (!rc) rc = check_cuda(cudaFuncSetAttribute(gae_flat4_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(G4Smem)), "gae_flat: smem attribute");
        if (rc) return rc;
        *configured = true;
    }
    static int prefetch_tiles = -1;
    if (prefetch_tiles < 0) {
        const char* env = getenv("G2048_GAE_PREFETCH");
        prefetch_tiles = env ? atoi(env) : sms / 2;
        if (prefetch_tiles < 0) prefetch_tiles = 0;
    }
    // persistent CTAs, all resident; a CTA holds two tiles, so no more CTAs than half the tiles (rounded up)
    int64_t grid = (int64_t)sms * G4_MIN_CTAS;
    const int64_t useful = (n_tiles + 1) / 2;
    if (grid > useful) grid = useful;
    cudaStream_t st = (cudaStream_t)stream;
    if (aligned) {
        gae_flat4_kernel<true><<<(unsigned)grid, G4_THREADS, sizeof(G4Smem), st>>>(
            d_rewards, d_values, d_dones, n, n_tiles, (float)gamma, (float)(gamma * lambda_gae), d_adv, d_ret,
            (G4Scratch*)d_scan_state, d_moments, prefetch_tiles);
    } else {
        gae_flat4_kernel<false><<<(unsigned)grid, G4_THREADS, sizeof(G4Smem), st>>>(
            d_rewards, d_values, d_dones, n, n_tiles, (float)gamma, (float)(gamma * lambda_gae), d_adv, d_ret,
            (G4Scratch*)d_scan_state, d_moments, prefetch_tiles);
    }
    G2048_CHECK_LAUNCH("gae_flat");
    if (d_moments) {
        gae_flat4_count_kernel<<<1, 1, 0, st>>>(d_moments, (double)n);
        G2048_CHECK_LAUNCH("gae_flat (count)");
    }
    return G2048_OK;
}

// g2048_gae_flat: the pipelined kernel from 2^23 steps on (a persistent CTA per two tiles needs enough tiles to fill
// the device), the one-tile-per-CTA kernel below that.  Measured on B200, episodes of ~300 steps, us:
//   steps        2^20   2^22   2^23   2^24   2^26
//   tiled        26.8   37.0   59.7   96.4  312
//   pipelined    33.1   40.4   58.4   85.3  260
#ifndef G2048_GAE4_NO_DISPATCH  // tools/probes/probe_gae4.cu builds this file alone
extern "C" int g2048_gae_flat_tiled(const float* d_rewards, const float* d_values, const uint8_t* d_dones, int64_t n,
                                    double gamma, double lambda_gae, float* d_adv, float* d_ret, void* d_scan_state,
                                    double* d_moments, void* stream);

extern "C" int g2048_gae_flat(const float* d_rewards, const float* d_values, const uint8_t* d_dones, int64_t n,
                              double gamma, double lambda_gae, float* d_adv, float* d_ret, void* d_scan_state,
                              double* d_moments, void* stream) {
    if (n >= (1ll << 23))
        return g2048_gae_flat_pipelined(d_rewards, d_values, d_dones, n, gamma, lambda_gae, d_adv, d_ret, d_scan_state, d_moments, stream);
    return g2048_gae_flat_tiled(d_rewards, d_values, d_dones, n, gamma, lambda_gae, d_adv, d_ret, d_scan_state, d_moments, stream);
}
#endif
