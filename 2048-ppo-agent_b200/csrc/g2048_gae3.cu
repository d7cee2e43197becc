// GAE over a flat buffer, current generation (src/ppo/data_loader.py:103-130; bit-identical results).
//
// What limits this kernel is not arithmetic but how many steps an SM keeps in flight while the serial
// recurrence of the longest episode of a tile runs (one multiply-add per step).  ncu on the earlier
// versions: v1 (1 024-step tiles, everything staged in shared memory) was issue-bound by one-lane
// walks; a 7 680-step version filled by bulk copies kept 9 bytes per step resident, so only 3 tiles
// fitted an SM and HBM idled while they walked (23 % DRAM throughput); a scalar-load version of the
// present layout spent 27 % of its samples waiting on 4-byte loads and 40 % at the barrier behind a
// walk that re-loaded shared memory every step.  Now:
//   phase 1  128-bit coalesced loads of r, V (done: 32-bit) straight into registers, a whole half
//            tile in flight per thread; delta = (r + gamma*V[t+1]) - V goes to shared memory (V[t+1]
//            of a lane's last step comes from the next lane by shuffle); four ballots of the done
//            flags, bit-interleaved, ARE the ordered episode list;
//   phase 2  one LANE per episode walks backwards over its deltas with 128-bit shared accesses,
//            software-pipelined two groups ahead; the tile's first episode (whose result the previous
//            tile waits for) and its open tail (which waits for the next tile) get warps of their own;
//   phase 3  advantages from shared memory, V again from global (an L2 hit: this CTA read it
//            microseconds ago), returns = adv + V, 128-bit streaming stores, fp64 moments.
// Only the 4 bytes per step that the walk needs live in shared memory (6 144 steps = 26 KiB per CTA).
// HBM traffic stays at the algorithmic 9 B read + 8 B written per step.
#include <cstdlib>

#include "g2048_common.cuh"
#include "g2048_gae_walk.cuh"

namespace g2048 {

// Geometry (overridable for tools/probes/probe_gae3.cu).  Throughput is (steps resident per SM) / (tile latency),
// and the latency is set by the longest episode of a tile, not by the tile's size: what counts is how many steps'
// deltas fit an SM's shared memory and that enough CTAs are resident to hold them.
#ifndef GAE3_THREADS_PER_CTA
#define GAE3_THREADS_PER_CTA 256
#endif
#ifndef GAE3_TILE_STEPS
#define GAE3_TILE_STEPS 6144
#endif
#ifndef GAE3_MIN_CTAS
#define GAE3_MIN_CTAS 6
#endif
constexpr int GAE3_THREADS = GAE3_THREADS_PER_CTA;
constexpr int GAE3_WARPS = GAE3_THREADS / 32;
constexpr int GAE3_TILE = GAE3_TILE_STEPS;
constexpr int GAE3_VEC_PER_THREAD = GAE3_TILE / (4 * GAE3_THREADS);  // float4 groups per thread
constexpr int GAE3_BLOCKS = GAE3_TILE / 128;                    // 128-step blocks: one warp-wide float4 access each
constexpr int GAE3_INFLIGHT = (GAE3_VEC_PER_THREAD % 3 == 0) ? 3 : 2;  // float4 groups a thread loads before it computes
constexpr int GAE3_PASSES = GAE3_VEC_PER_THREAD / GAE3_INFLIGHT;
static_assert(GAE3_WARPS == 4 || GAE3_WARPS == 8, "four or eight warps per CTA");
static_assert(GAE3_VEC_PER_THREAD * 4 * GAE3_THREADS == GAE3_TILE && GAE3_PASSES * GAE3_INFLIGHT == GAE3_VEC_PER_THREAD, "tile shape");
// warp roles in phase 2 (rotated over the warp schedulers with the ticket)
constexpr int GAE3_ROLE_FIRST = GAE3_WARPS == 4 ? 1 : 4, GAE3_ROLE_TAIL = GAE3_WARPS == 4 ? 2 : 5;
constexpr int GAE3_WALK_SLOTS = GAE3_WARPS - 2;

#ifdef G2048_GAE_TIMELINE  // tools/probes/probe_gae3.cu only: per-CTA phase timestamps
__device__ long long g_gae_timeline[16 * 65536];
#define GAE3_STAMP(k) do { if (threadIdx.x == 0 && blockIdx.x < 65536) g_gae_timeline[16 * blockIdx.x + (k)] = clock64(); } while (0)
#define GAE3_STAMP_ANY(k) do { if (blockIdx.x < 65536) g_gae_timeline[16 * blockIdx.x + (k)] = clock64(); } while (0)
#else
#define GAE3_STAMP(k) do { } while (0)
#define GAE3_STAMP_ANY(k) do { } while (0)
#endif

struct Gae3Scratch {
    unsigned int ticket;
    unsigned int pad[3];
};

struct Gae3Smem {
    float g[GAE3_TILE];                 // delta -> advantages, in place
    uint32_t ballot[4 * GAE3_BLOCKS];   // per 128-step block: ballot j has bit l set iff step 128*b + 4*l + j is a done
    uint32_t pref[GAE3_BLOCKS + 1];     // exclusive prefix of the blocks' done counts
    double red[4 * (GAE3_THREADS / 32)];
    unsigned int ticket;
};

// position (in the tile) of the k-th done step, k < n_done
__device__ __forceinline__ int gae3_locate(const Gae3Smem& s, int k) { return gae_locate(s.ballot, s.pref, GAE3_BLOCKS, k); }

template <bool ALIGNED>
__global__ void __launch_bounds__(GAE3_THREADS, GAE3_MIN_CTAS)
gae_flat3_kernel(const float* __restrict__ rewards, const float* __restrict__ values, const uint8_t* __restrict__ dones,
                 int64_t n, int64_t n_tiles, float gamma, float gamma_lambda, float* __restrict__ adv,
                 float* __restrict__ ret, Gae3Scratch* scratch, double* __restrict__ moments, int prefetch_tiles) {
    __shared__ __align__(16) Gae3Smem s;
    volatile unsigned int* flags = (volatile unsigned int*)(scratch + 1);
    volatile float* heads = (volatile float*)((unsigned int*)(scratch + 1) + n_tiles);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    GAE3_STAMP(0);
    if (tid == 0) {
        race_jitter();
        s.ticket = atomicAdd(&scratch->ticket, 1u);
    }
    __syncthreads();
    const int64_t tile = n_tiles - 1 - (int64_t)s.ticket;  // tiles are taken from the END of the buffer
    const int64_t lo = tile * GAE3_TILE;
    const int len = (int)min((int64_t)GAE3_TILE, n - lo);
    GAE3_STAMP(1);
    // Only the CTAs that are in phase 1 have loads in flight, too few bytes to keep HBM busy.  So the inputs of the
    // tile that will be taken `prefetch_tiles` tickets from now (about when this CTA's SM slot frees up) are pulled
    // into L2 now: phase 1 then runs at L2 latency and DRAM sees a steady stream that no CTA waits for.
    if (prefetch_tiles > 0 && warp == GAE3_WARPS - 1 && tile - prefetch_tiles >= 0)  // a full tile if it exists
        gae_tile_prefetch<GAE3_TILE>(rewards, values, dones, tile - prefetch_tiles, lane);

    // ---- phase 1: delta into shared memory, done bits into ballots ---------------------------------------------
    gae_tile_deltas<ALIGNED, GAE3_TILE, GAE3_THREADS, GAE3_INFLIGHT>(s.g, s.ballot, rewards, values, dones, n, lo, len, gamma, tid);
    __syncthreads();
    GAE3_STAMP(2);
    if (warp == 0) gae_tile_prefix<GAE3_BLOCKS>(s.ballot, s.pref, lane);
    __syncthreads();
    const int n_done = (int)s.pref[GAE3_BLOCKS];
    GAE3_STAMP(3);

    // ---- phase 2: one lane per episode ---------------------------------------------------------------------
    // Warp w issues on scheduler w % 4.  The walking warps of the CTAs that share an SM must not pile up
    // on one scheduler (timeline probe: with fixed roles the five "warp 0" walkers of an SM shared
    // scheduler 0 and the walk ran at 35 cycles per step), so the roles rotate with the ticket.
    const int rot = (int)(s.ticket & 3u);
    // eight warps: roles 0-3, 6-7 walk, 4 first episode, 5 tail; four warps: 0 and 3 walk, 1 first episode, 2 tail
    const int role = (warp < 4) ? ((warp - rot) & 3) : 4 + ((warp - rot) & 3);
    if (role == GAE3_ROLE_FIRST) {
        // the first episode of the tile alone in its warp: the previous tile is waiting for its result
        if (lane == 0 && n_done > 0) {
#ifndef G2048_GAE3_SKIP_SIDE_WALKS  // timing experiment only (tools/probes/probe_gae3.cu)
            gae_walk(s.g, gae3_locate(s, 0), -1, 0.0f, gamma_lambda);
#endif
            race_jitter();
            heads[tile] = s.g[0];
            __threadfence();
            flags[tile] = 1u;
            GAE3_STAMP_ANY(8);   // first episode published
        }
    } else if (role == GAE3_ROLE_TAIL) {
        // the steps after the tile's last done belong to an episode that ends in a later tile
        if (lane == 31) {
            const int first_excl = n_done ? gae3_locate(s, n_done - 1) : -1;
            if (first_excl < len - 1) {
                float carry = 0.0f;
                if (lo + len < n) {
                    race_jitter();
                    while (flags[tile + 1] == 0u) __nanosleep(100);
                    __threadfence();
                    carry = heads[tile + 1];
                }
                GAE3_STAMP_ANY(9);   // look-back satisfied
#ifndef G2048_GAE3_SKIP_SIDE_WALKS
                gae_walk(s.g, len - 1, first_excl, carry, gamma_lambda);
#endif
                GAE3_STAMP_ANY(10);  // tail walked
            }
            if (n_done == 0) {
                race_jitter();
            heads[tile] = s.g[0];
                __threadfence();
                flags[tile] = 1u;
            }
        }
    } else {
        // episodes 1.. : the first 32 on the first walking warp, the next 32 on the second, ...
        const int slot = GAE3_WARPS == 4 ? (role == 0 ? 0 : 1) : (role < 4 ? role : role - 2);
        for (int base = 1 + 32 * slot; base < n_done; base += GAE3_WALK_SLOTS * 32) {  // warp-uniform
            const int e = base + lane;
            const int end = e < n_done ? gae3_locate(s, e) : 0;
            int prev = __shfl_up_sync(0xFFFFFFFFu, end, 1);  // the episode before mine ends where my neighbour's does
            if (lane == 0) prev = gae3_locate(s, base - 1);
#ifndef G2048_GAE3_SKIP_MAIN_WALKS  // timing experiment only
            if (e < n_done) gae_walk(s.g, end, prev, 0.0f, gamma_lambda);
#endif
        }
        if (slot == 0 && lane == 0) GAE3_STAMP_ANY(11);  // the first walking warp finished its episodes
    }
    __syncthreads();
    GAE3_STAMP(4);

    // ---- phase 3: returns, stores, moments -------------------------------------------------------------------
    double m[4] = {0.0, 0.0, 0.0, 0.0};
    gae_tile_store<ALIGNED, GAE3_TILE, GAE3_THREADS, GAE3_INFLIGHT>(s.g, values, lo, len, adv, ret, m, tid);
    GAE3_STAMP(5);
    if (moments) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double x = m[k];
            for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xFFFFFFFFu, x, off);
            if (lane == 0) s.red[k * (GAE3_THREADS / 32) + warp] = x;
        }
        __syncthreads();
        if (tid < 4) {
            double t = 0.0;
            for (int w = 0; w < GAE3_THREADS / 32; ++w) t += s.red[tid * (GAE3_THREADS / 32) + w];
            atomicAdd(&moments[1 + tid], t);
        }
        if (tid == 4) atomicAdd(&moments[0], (double)len);
    }
    GAE3_STAMP(6);
}

}  // namespace g2048

using namespace g2048;

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

extern "C" int g2048_gae_flat_tiled(const float* d_rewards, const float* d_values, const uint8_t* d_dones, int64_t n,
                              double gamma, double lambda_gae, float* d_adv, float* d_ret, void* d_scan_state,
                              double* d_moments, void* stream) {
    G2048_REQUIRE(n >= 0, "gae_flat: n");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_rewards && d_values && d_dones && d_adv && d_ret && d_scan_state, "gae_flat: pointers");
    const int64_t n_tiles = (n + GAE3_TILE - 1) / GAE3_TILE;
    // 128-bit accesses need 16-byte aligned float arrays and a 4-byte aligned done array
    const bool aligned = aligned16(d_rewards) && aligned16(d_values) && aligned16(d_adv) && aligned16(d_ret) &&
                         ((uintptr_t)d_dones & 3u) == 0;
    static bool configured_on[64] = {false};
    bool* configured = device_once_flag(configured_on);
    if (!configured) return fail_arg("no CUDA device");
    if (!*configured) {  // several CTAs of ~26 KiB per SM need the shared-memory-heavy L1 split
        cudaFuncSetAttribute(gae_flat3_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(gae_flat3_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        *configured = true;
    }
    cudaStream_t st = (cudaStream_t)stream;
    // prefetch distance in tiles (G2048_GAE_PREFETCH overrides, 0 = off).  Swept on B200 with tools/probes/probe_gae3.cu
    // at 2^26 steps: 0 -> 334 us, SMs/2 = 74 -> 312 us, 148 -> 315, 296 -> 320, 444 -> 338, 888 -> 372 (too early:
    // the lines are gone again before their tile starts)
    static int prefetch_tiles = -1;
    if (prefetch_tiles < 0) {
        const char* env = getenv("G2048_GAE_PREFETCH");
        const int sms = sm_count();
        prefetch_tiles = env ? atoi(env) : (sms > 0 ? sms / 2 : 0);
        if (prefetch_tiles < 0) prefetch_tiles = 0;
    }
    if (aligned) {
        gae_flat3_kernel<true><<<(unsigned)n_tiles, GAE3_THREADS, 0, st>>>(
            d_rewards, d_values, d_dones, n, n_tiles, (float)gamma, (float)(gamma * lambda_gae), d_adv, d_ret,
            (Gae3Scratch*)d_scan_state, d_moments, prefetch_tiles);
    } else {
        gae_flat3_kernel<false><<<(unsigned)n_tiles, GAE3_THREADS, 0, st>>>(
            d_rewards, d_values, d_dones, n, n_tiles, (float)gamma, (float)(gamma * lambda_gae), d_adv, d_ret,
            (Gae3Scratch*)d_scan_state, d_moments, prefetch_tiles);
    }
    G2048_CHECK_LAUNCH("gae_flat");
    return G2048_OK;
}
