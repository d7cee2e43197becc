// GAE / returns and their normalisation (src/ppo/data_loader.py:103-130 and :61-67).
//
// The reference walks the flat buffer backwards one step at a time.  The recurrence only couples
// steps of the same episode (a `done` resets the carry), so the kernels parallelise ACROSS
// episodes and keep the reference's exact fp32 operation order INSIDE each episode:
//     delta = (r[t] + gamma * last_v) - V[t];   gae = delta + (gamma*lambda) * gae
// (no FMA contraction: the library is built with --fmad=false).  Results are therefore bit-identical
// to the reference, not merely within tolerance.  The kernels are HBM-bound streams:
// 9 B read + 8 B written per step.
#include "g2048_common.cuh"
#include "g2048_hostcopy.cuh"
#include "g2048_tma.cuh"

namespace g2048 {

#ifndef G2048_GAE_EXPERIMENT
#define G2048_GAE_EXPERIMENT 0  // non-zero only in tools/probes/probe_gae.cu
#endif

constexpr int GAE_TILE = 1024;
#ifdef G2048_LEGACY_KERNELS
constexpr int GAE_THREADS = 128;
constexpr int GAE_ITEMS = GAE_TILE / GAE_THREADS;  // 8 consecutive steps per thread for segment discovery
#endif

struct GaeScratch {
    unsigned int ticket;
    unsigned int pad[3];
    // followed by n_tiles {flag, head_gae} pairs
};

__device__ __forceinline__ void block_sum4(double v[4], double* s_red /* [4][4] */, double* __restrict__ moments) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        double x = v[k];
        for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xFFFFFFFFu, x, off);
        if (lane == 0) s_red[k * 4 + warp] = x;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        const int k = threadIdx.x;
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[k * 4 + w];
        atomicAdd(&moments[1 + k], t);
    }
}

#ifdef G2048_LEGACY_KERNELS  // first-generation GAE kernel: built only into the tests' libg2048_legacy.so
// One CTA per tile of 1024 consecutive steps, tiles taken from the END of the buffer by ticket.
//   1. coalesced loads of r, V, done into shared memory (+ V of the step after the tile);
//   2. warp-shuffle prefix sum of per-thread done counts -> ordered list of episode ends;
//   3. delta[t] for every step in parallel; then one thread per episode end walks its episode
//      backwards inside shared memory with one multiply-add per step;
//   4. the steps after the tile's last `done` belong to an episode that ends in a later tile:
//      one thread waits for that tile's first-step gae (published through global memory, decoupled
//      look-back of depth one) and walks them;
//   5. coalesced stores of adv / ret, fp64 moment sums for the normalisation.
__global__ void __launch_bounds__(GAE_THREADS)
gae_flat_kernel(const float* __restrict__ rewards, const float* __restrict__ values, const uint8_t* __restrict__ dones,
                int64_t n, int64_t n_tiles, float gamma, float gamma_lambda, float* __restrict__ adv,
                float* __restrict__ ret, GaeScratch* scratch, double* __restrict__ moments) {
    __shared__ float s_r[GAE_TILE];
    __shared__ float s_v[GAE_TILE + 1];
    __shared__ float s_adv[GAE_TILE];
    __shared__ uint8_t s_d[GAE_TILE];
    __shared__ unsigned short s_end[GAE_TILE];  // ordered positions of done steps
    __shared__ int s_warp_count[GAE_THREADS / 32];
    __shared__ double s_red[16];
    __shared__ unsigned int s_ticket;

    volatile unsigned int* flags = (volatile unsigned int*)(scratch + 1);
    volatile float* heads = (volatile float*)((unsigned int*)(scratch + 1) + n_tiles);

    if (threadIdx.x == 0) s_ticket = atomicAdd(&scratch->ticket, 1u);
    __syncthreads();
    const int64_t tile = n_tiles - 1 - (int64_t)s_ticket;  // memory-order index of this tile
    const int64_t lo = tile * GAE_TILE;
    const int len = (int)min((int64_t)GAE_TILE, n - lo);

    for (int i = threadIdx.x; i < GAE_TILE; i += GAE_THREADS) {
        const bool in = i < len;
        s_r[i] = in ? rewards[lo + i] : 0.0f;
        s_v[i] = in ? values[lo + i] : 0.0f;
        s_d[i] = in ? dones[lo + i] : (uint8_t)0;
    }
    if (threadIdx.x == 0) s_v[len] = (lo + len < n) ? values[lo + len] : 0.0f;  // V of the next step (0 past the end)
    __syncthreads();

    // delta[t] = (r[t] + gamma * V[t+1]) - V[t], with V[t+1] := 0 after a done step -- elementwise, so
    // every thread helps; the serial part below is then one multiply-add per step.
    float delta_reg[GAE_ITEMS];
#pragma unroll
    for (int k = 0; k < GAE_ITEMS; ++k) {
        const int i = threadIdx.x + k * GAE_THREADS;
        const float next_v = s_d[i] ? 0.0f : s_v[i + 1];
        delta_reg[k] = (s_r[i] + gamma * next_v) - s_v[i];
    }

    // ordered list of done positions
    const int base = threadIdx.x * GAE_ITEMS;
    unsigned int bits = 0;
#pragma unroll
    for (int k = 0; k < GAE_ITEMS; ++k) bits |= (s_d[base + k] ? 1u : 0u) << k;
    const int cnt = __popc(bits);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = cnt;
    for (int off = 1; off < 32; off <<= 1) {
        const int y = __shfl_up_sync(0xFFFFFFFFu, incl, off);
        if (lane >= off) incl += y;
    }
    if (lane == 31) s_warp_count[warp] = incl;
    __syncthreads();  // also: every s_r[i] has been read into delta_reg
#pragma unroll
    for (int k = 0; k < GAE_ITEMS; ++k) s_r[threadIdx.x + k * GAE_THREADS] = delta_reg[k];
    int before = incl - cnt;
    int n_done = 0;
#pragma unroll
    for (int w = 0; w < GAE_THREADS / 32; ++w) {
        if (w < warp) before += s_warp_count[w];
        n_done += s_warp_count[w];
    }
    {
        unsigned int b = bits;
        int slot = before;
        while (b) {
            const int k = __ffs(b) - 1;
            s_end[slot++] = (unsigned short)(base + k);
            b &= b - 1;
        }
    }
    __syncthreads();

    // walk one episode (or fragment) backwards over steps (first_excl, last]: gae = delta + gl * gae.
    // A done step starts with gae = 0, which is what the reference's reset computes (delta + gl * 0).
    auto walk = [&](int last, int first_excl, float g) {
        int t = last;
#if G2048_GAE_EXPERIMENT & 1  // timing experiment (tools/probes/probe_gae.cu): no serial chain
        s_adv[last] = g;
        return;
#endif
        for (; t - 8 >= first_excl; t -= 8) {
            float d[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) d[k] = s_r[t - k];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                g = d[k] + gamma_lambda * g;
                s_adv[t - k] = g;
            }
        }
        for (; t > first_excl; --t) {
            g = s_r[t] + gamma_lambda * g;
            s_adv[t] = g;
        }
    };

    // episodes that end inside the tile; episode 0 also fixes the tile's first-step gae
    for (int s = threadIdx.x; s < n_done; s += GAE_THREADS) {
        const int last = s_end[s];
        const int first_excl = s ? (int)s_end[s - 1] : -1;
        walk(last, first_excl, 0.0f);
        if (s == 0) {
            heads[tile] = s_adv[0];
            __threadfence();
            flags[tile] = 1u;
        }
    }
    // trailing fragment: continues an episode that ends in a later tile (or is cut by the buffer end)
    if (threadIdx.x == GAE_THREADS - 1) {
        const int first_excl = n_done ? (int)s_end[n_done - 1] : -1;
        if (first_excl < len - 1) {
            float carry = 0.0f;
            if (lo + len < n) {
#if !(G2048_GAE_EXPERIMENT & 4)
                while (flags[tile + 1] == 0u) __nanosleep(32);
#endif
                __threadfence();
                carry = heads[tile + 1];
            }
            walk(len - 1, first_excl, carry);
        }
        if (n_done == 0) {
            heads[tile] = s_adv[0];
            __threadfence();
            flags[tile] = 1u;
        }
    }
    __syncthreads();

    double m[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < len; i += GAE_THREADS) {
        const float a = s_adv[i];
        const float rt = a + s_v[i];
        adv[lo + i] = a;
        ret[lo + i] = rt;
        m[0] += (double)a;
        m[1] += (double)a * (double)a;
        m[2] += (double)rt;
        m[3] += (double)rt * (double)rt;
    }
    if (moments) {
        block_sum4(m, s_red, moments);
        if (threadIdx.x == 0) atomicAdd(&moments[0], (double)len);
    }
}
#endif  // G2048_LEGACY_KERNELS

// Time-major (T,B) records: one lane per env, loads batched 16 steps ahead so that enough bytes
// are in flight even at B = 64 K (one lane per env is all the parallelism there is); coalesced along B.
// (A variant with two register sets -- next 16 steps' loads in flight during the current 16 steps' stores -- needed
// 150 registers and was slower: 45 vs 36 us for 128 x 65 536 steps, tools/probes/tm_bench.py; 4.9 TB/s at 128 x 262 144.
// Unroll 8: 45 us, 16: 36 us, 32: 43 us.  ncu then showed 56 instructions per step, a third of them predicates and
// 64-bit index arithmetic: with a predicate-free path for full batches, row pointers stepped by n and fma for the
// squared moments 36 -> 33 us (94 -> 64 registers); unroll 8 / 32 and 32-thread CTAs re-measured: 36 / 37 / 37 us.)
#ifndef G2048_TM_UNROLL
#define G2048_TM_UNROLL 16
#endif
#ifndef G2048_TM_THREADS
#define G2048_TM_THREADS 64
#endif
#ifndef G2048_TM_MIN_CTAS
#define G2048_TM_MIN_CTAS 1
#endif
constexpr int GAE_TM_UNROLL = G2048_TM_UNROLL;
constexpr int GAE_TM_THREADS = G2048_TM_THREADS;

__global__ void __launch_bounds__(GAE_TM_THREADS, G2048_TM_MIN_CTAS)
gae_time_major_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                      const uint8_t* __restrict__ meta, int64_t t_steps, int64_t n,
                      const float* __restrict__ bootstrap, float gamma, float gamma_lambda, float* __restrict__ adv,
                      float* __restrict__ ret, double* __restrict__ moments) {
    __shared__ double s_red[16];
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double m[4] = {0.0, 0.0, 0.0, 0.0};
    if (e < n) {
        float last_v = bootstrap ? bootstrap[e] : 0.0f;
        float last_gae = 0.0f;
        // one step of the reference's loop (data_loader.py:110-128) in its fp32 order; the squares go through fma --
        // the product of two floats is exact in double, so mul + add and fma give the same sums
        auto step = [&](float r, float v, uint32_t d, float* __restrict__ pa, float* __restrict__ pr) {
            if (d & 0x40u) {
                last_v = 0.0f;
                last_gae = 0.0f;
            }
            const float delta = (r + gamma * last_v) - v;
            last_gae = delta + gamma_lambda * last_gae;
            const float rt = last_gae + v;
            __stcs(pa, last_gae);
            __stcs(pr, rt);
            last_v = v;
            const double da = (double)last_gae, dr = (double)rt;
            m[0] += da;
            m[1] = __fma_rn(da, da, m[1]);
            m[2] += dr;
            m[3] = __fma_rn(dr, dr, m[3]);
        };
        int64_t t = t_steps - 1;
        // full batches: no per-step bounds test, one row pointer per array stepped back by n (ncu on the previous
        // form: 56 instructions per step, a third of them predicates and 64-bit index arithmetic)
        while (t >= GAE_TM_UNROLL - 1) {
            const int64_t row = t * n + e;
            const float* pr = rewards + row;
            const float* pv = values + row;
            const uint8_t* pd = meta + row;
            float r[GAE_TM_UNROLL], v[GAE_TM_UNROLL];
            uint32_t d[GAE_TM_UNROLL];
#pragma unroll
            for (int k = 0; k < GAE_TM_UNROLL; ++k) {
                r[k] = __ldcs(pr);
                v[k] = __ldcs(pv);
                d[k] = __ldcs(pd);
                pr -= n;
                pv -= n;
                pd -= n;
            }
            float* pa = adv + row;
            float* pt = ret + row;
#pragma unroll
            for (int k = 0; k < GAE_TM_UNROLL; ++k) {
                step(r[k], v[k], d[k], pa, pt);
                pa -= n;
                pt -= n;
            }
            t -= GAE_TM_UNROLL;
        }
        if (t >= 0) {  // fewer than a batch left (the oldest steps): the same batch, predicated
            const int cnt = (int)t + 1;
            float r[GAE_TM_UNROLL], v[GAE_TM_UNROLL];
            uint32_t d[GAE_TM_UNROLL];
#pragma unroll
            for (int k = 0; k < GAE_TM_UNROLL - 1; ++k) {
                if (k < cnt) {
                    const int64_t i = (t - k) * n + e;
                    r[k] = __ldcs(&rewards[i]);
                    v[k] = __ldcs(&values[i]);
                    d[k] = __ldcs(&meta[i]);
                }
            }
#pragma unroll
            for (int k = 0; k < GAE_TM_UNROLL - 1; ++k) {
                if (k < cnt) {
                    const int64_t i = (t - k) * n + e;
                    step(r[k], v[k], d[k], adv + i, ret + i);
                }
            }
        }
    }
    if (moments) {
        block_sum4(m, s_red, moments);
        if (threadIdx.x == 0) {
            const int64_t first = (int64_t)blockIdx.x * blockDim.x;
            const int64_t envs = min((int64_t)blockDim.x, n - first);
            atomicAdd(&moments[0], (double)(envs * t_steps));
        }
    }
}

// The same recurrence fed through a bulk-copy ring (round 2, last form).  What bounds the kernel above at C3 is not
// bandwidth but phases: a lane issues 16 steps' loads, waits a DRAM round trip, computes, stores, and only then asks
// for the next 16 -- all 14 warps of an SM in step with each other -- so DRAM sees bursts (ncu: 23 us of kernel for
// 95 MB, the launch's own 47 MB of dirty lines written back behind it).  Here a CTA of 128 lanes owns 128 envs and
// its first warp keeps THREE stages of 16 rows (rewards, values, meta: 18 KiB a stage) in flight through cp.async.bulk,
// one 512- or 128-byte row per copy (a lane per row), completion on an mbarrier per stage; the lanes walk a stage from shared memory while
// the copy engine fills the other two, so loads, arithmetic and stores overlap without a register for any of it.
// Same fp32 operation order per env: bit-identical.  Needs n % 128 == 0 and 16-byte aligned rows (else the kernel above).
// Measured (tools/probes/tm_ring_probe.py, C3 / 4 x C3, steady state): one lane per env 33.7 / 110 us; ring with 64-env
// CTAs 33.6 / 106; 128-env CTAs 29.5 / 103 (longer rows: fewer DRAM pages per byte); 32 rows a stage 43 / 133.
#ifndef G2048_TM_RING_ROWS
#define G2048_TM_RING_ROWS 16
#endif
#ifndef G2048_TM_RING_ENVS
#define G2048_TM_RING_ENVS 128
#endif
constexpr int GAE_RING_ROWS = G2048_TM_RING_ROWS;
constexpr int GAE_RING_STAGES = 3;
constexpr int GAE_RING_ENVS = G2048_TM_RING_ENVS;
static_assert(GAE_RING_ENVS % 32 == 0 && GAE_RING_ENVS <= 128, "block_sum4 reduces at most four warps");
struct GaeRingStage {
    float r[GAE_RING_ROWS][GAE_RING_ENVS];
    float v[GAE_RING_ROWS][GAE_RING_ENVS];
    uint8_t d[GAE_RING_ROWS][GAE_RING_ENVS];
};

__global__ void __launch_bounds__(GAE_RING_ENVS)
gae_time_major_ring_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                           const uint8_t* __restrict__ meta, int64_t t_steps, int64_t n,
                           const float* __restrict__ bootstrap, float gamma, float gamma_lambda, float* __restrict__ adv,
                           float* __restrict__ ret, double* __restrict__ moments) {
    extern __shared__ __align__(128) uint8_t ring_smem[];
    GaeRingStage* stages = reinterpret_cast<GaeRingStage*>(ring_smem);
    __shared__ __align__(8) uint64_t s_full[GAE_RING_STAGES];
    __shared__ double s_red[16];
    const int64_t e0 = (int64_t)blockIdx.x * GAE_RING_ENVS;
    const int64_t e = e0 + threadIdx.x;
    const int64_t n_batches = (t_steps + GAE_RING_ROWS - 1) / GAE_RING_ROWS;
    if (threadIdx.x == 0) {
        for (int k = 0; k < GAE_RING_STAGES; ++k) mbar_init(&s_full[k], 1);
        fence_proxy_async();  // the initialised barriers -> visible to the copy engine
    }
    __syncthreads();
    // batch b (0 = the newest steps) holds rows [lo, hi), hi = T - 16 b; row t sits at index t - lo of its stage
    auto issue = [&](int64_t b) {  // warp 0: lane 0 announces the stage's bytes, then one row per lane
        if (b >= n_batches) return;
        const int64_t hi = t_steps - b * GAE_RING_ROWS, lo = hi > GAE_RING_ROWS ? hi - GAE_RING_ROWS : 0;
        const int stage = (int)(b % GAE_RING_STAGES);
        GaeRingStage& st = stages[stage];
        if (threadIdx.x == 0) mbar_arrive_expect_tx(&s_full[stage], (uint32_t)((hi - lo) * (GAE_RING_ENVS * 9)));
        __syncwarp();
        for (int64_t t = lo + threadIdx.x; t < hi; t += 32) {
            const int64_t row = t * n + e0;
            bulk_load(st.r[t - lo], rewards + row, GAE_RING_ENVS * 4, &s_full[stage]);
            bulk_load(st.v[t - lo], values + row, GAE_RING_ENVS * 4, &s_full[stage]);
            bulk_load(st.d[t - lo], meta + row, GAE_RING_ENVS, &s_full[stage]);
        }
    };
    if (threadIdx.x < 32) {
        issue(0);
        issue(1);
    }
    double m[4] = {0.0, 0.0, 0.0, 0.0};
    float last_v = bootstrap ? bootstrap[e] : 0.0f;
    float last_gae = 0.0f;
    for (int64_t b = 0; b < n_batches; ++b) {
        // two batches ahead: that stage was last read in batch b - 1, before the barrier that ended it
        if (threadIdx.x < 32) issue(b + 2);
        const int64_t hi = t_steps - b * GAE_RING_ROWS, lo = hi > GAE_RING_ROWS ? hi - GAE_RING_ROWS : 0;
        const int stage = (int)(b % GAE_RING_STAGES);
        const GaeRingStage& st = stages[stage];
        mbar_wait(&s_full[stage], (uint32_t)((b / GAE_RING_STAGES) & 1));
        float* pa = adv + (hi - 1) * n + e;
        float* pr = ret + (hi - 1) * n + e;
        for (int k = (int)(hi - lo) - 1; k >= 0; --k) {
            // one step of the reference's loop (data_loader.py:110-128) in its fp32 order, as in gae_time_major_kernel
            const float r = st.r[k][threadIdx.x], v = st.v[k][threadIdx.x];
            if (st.d[k][threadIdx.x] & 0x40u) {
                last_v = 0.0f;
                last_gae = 0.0f;
            }
            const float delta = (r + gamma * last_v) - v;
            last_gae = delta + gamma_lambda * last_gae;
            const float rt = last_gae + v;
            __stcs(pa, last_gae);
            __stcs(pr, rt);
            last_v = v;
            const double da = (double)last_gae, dr = (double)rt;
            m[0] += da;
            m[1] = __fma_rn(da, da, m[1]);
            m[2] += dr;
            m[3] = __fma_rn(dr, dr, m[3]);
            pa -= n;
            pr -= n;
        }
        __syncthreads();  // every lane is done with this stage: warp 0 may refill it at the top of the next batch
    }
    if (moments) {
        block_sum4(m, s_red, moments);
        if (threadIdx.x == 0) atomicAdd(&moments[0], (double)(GAE_RING_ENVS * t_steps));
    }
}

// The same recurrence with the LOADS of a CTA's envs spread over four warps (round 2).  One lane per env leaves 13 warps
// per SM at C3 (65 536 envs), each walking 128 steps with 16 loads in flight: 0.64 of HBM.  Here a CTA is 32 envs x 4
// warps: all four fetch a 128-step chunk of the CTA's envs into shared memory (36 KiB: every load of the chunk is in
// flight at once), then warp 0 runs the recurrence over it from shared memory -- the reference's exact serial fp32
// order per env, bit-identical like the kernel above -- while warps 1-3 already fetch the chunk before it (T > 128).
// NOT the default: it measured slower than one lane per env (C3: 46 vs 34 us; 4 x C3: 150 vs 113 us) -- the recurrence
// of 32 envs on ONE warp per CTA, a shared-memory load and two global stores inside every dependent step, outweighs the
// wider load front.  Kept behind G2048_GAE_TM_SPLIT=1 for the record of the experiment.
// (First attempt: the chunk in REGISTERS, 32 steps per warp, the carry handed from warp to warp.  With 64 data
// registers per thread ptxas sank every load to just before its use: 128 serial DRAM round trips per env, 168 us for C3
// against 34 -- ncu long_scoreboard 15 per issue, profiles/r02_gae_kernels.csv.)
constexpr int GAE_TS_CHUNK = 128;
constexpr int GAE_TS_WARPS = 4;
struct GaeTsStage {
    float r[GAE_TS_CHUNK][32];
    float v[GAE_TS_CHUNK][32];
    uint8_t d[GAE_TS_CHUNK][32];
};

__global__ void __launch_bounds__(32 * GAE_TS_WARPS)
gae_time_major_split_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                            const uint8_t* __restrict__ meta, int64_t t_steps, int64_t n,
                            const float* __restrict__ bootstrap, float gamma, float gamma_lambda, float* __restrict__ adv,
                            float* __restrict__ ret, double* __restrict__ moments) {
    extern __shared__ __align__(16) uint8_t ts_smem[];
    GaeTsStage* stages = reinterpret_cast<GaeTsStage*>(ts_smem);  // one stage, or two when there is more than one chunk
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t e = (int64_t)blockIdx.x * 32 + lane;
    const bool live = e < n;
    const int64_t n_chunks = (t_steps + GAE_TS_CHUNK - 1) / GAE_TS_CHUNK;

    // rows [c * 128, ...) of my env into stage `st`, by `nl` loader warps of which I am number `lw`
    auto load_chunk = [&](int64_t c, GaeTsStage& st, int lw, int nl) {
        const int64_t t0 = c * GAE_TS_CHUNK;
        const int rows = (int)min((int64_t)GAE_TS_CHUNK, t_steps - t0);
        if (!live) return;
#pragma unroll 8
        for (int j = lw; j < rows; j += nl) {
            const int64_t i = (t0 + j) * n + e;
            st.r[j][lane] = __ldcs(&rewards[i]);
            st.v[j][lane] = __ldcs(&values[i]);
            st.d[j][lane] = __ldcs(&meta[i]);
        }
    };
    load_chunk(n_chunks - 1, stages[(n_chunks - 1) & 1], warp, GAE_TS_WARPS);
    __syncthreads();

    float last_v = (live && bootstrap) ? bootstrap[e] : 0.0f, last_gae = 0.0f;  // warp 0's carry
    double m[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t c = n_chunks - 1; c >= 0; --c) {
        if (warp == 0) {
            if (live) {
                const GaeTsStage& st = stages[c & 1];
                const int64_t t0 = c * GAE_TS_CHUNK;
                const int rows = (int)min((int64_t)GAE_TS_CHUNK, t_steps - t0);
                for (int j = rows - 1; j >= 0; --j) {
                    const float r = st.r[j][lane], v = st.v[j][lane];
                    if (st.d[j][lane] & 0x40u) {
                        last_v = 0.0f;
                        last_gae = 0.0f;
                    }
                    const float delta = (r + gamma * last_v) - v;
                    last_gae = delta + gamma_lambda * last_gae;
                    const float rt = last_gae + v;
                    const int64_t i = (t0 + j) * n + e;
                    __stcs(&adv[i], last_gae);
                    __stcs(&ret[i], rt);
                    last_v = v;
                    const double da = (double)last_gae, dr = (double)rt;
                    m[0] += da;
                    m[1] = __fma_rn(da, da, m[1]);
                    m[2] += dr;
                    m[3] = __fma_rn(dr, dr, m[3]);
                }
            }
        } else if (c > 0) {
            load_chunk(c - 1, stages[(c - 1) & 1], warp - 1, GAE_TS_WARPS - 1);  // the chunk before this one, meanwhile
        }
        __syncthreads();
    }
    if (moments && warp == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double x = m[k];
            for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xFFFFFFFFu, x, off);
            if (lane == 0) atomicAdd(&moments[1 + k], x);
        }
        if (lane == 0) {
            const int64_t envs = min((int64_t)32, n - (int64_t)blockIdx.x * 32);
            atomicAdd(&moments[0], (double)(envs * t_steps));
        }
    }
}

// x = (x - mean) / (std_unbiased + 1e-8), 128-bit accesses on the aligned body
__global__ void __launch_bounds__(256)
normalize_kernel(float* __restrict__ x, int64_t n, const double* __restrict__ moments, int which) {
    const double cnt = moments[0];
    const double sum = moments[which], sumsq = moments[which + 1];
    const double mean_d = sum / cnt;
    const double var_d = (sumsq - sum * mean_d) / (cnt - 1.0);
    const float mean = (float)mean_d;
    const float denom = (float)sqrt(var_d > 0.0 ? var_d : 0.0) + 1e-8f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // head up to 16-byte alignment, vector body, tail
    const int64_t head = min(n, (int64_t)((16 - ((uintptr_t)x & 15u)) & 15u) / 4);
    for (int64_t i = tid; i < head; i += stride) x[i] = (x[i] - mean) / denom;
    float4* xv = (float4*)(x + head);
    const int64_t nv = (n - head) / 4;
    for (int64_t i = tid; i < nv; i += stride) {
        float4 v = xv[i];
        v.x = (v.x - mean) / denom;
        v.y = (v.y - mean) / denom;
        v.z = (v.z - mean) / denom;
        v.w = (v.w - mean) / denom;
        xv[i] = v;
    }
    for (int64_t i = head + nv * 4 + tid; i < n; i += stride) x[i] = (x[i] - mean) / denom;
}

}  // namespace g2048

using namespace g2048;

extern "C" int64_t g2048_gae_flat_scratch_bytes(int64_t n) {
    const int64_t n_tiles = (n + GAE_TILE - 1) / GAE_TILE;
    return (int64_t)sizeof(GaeScratch) + n_tiles * 8;
}

#ifdef G2048_LEGACY_KERNELS
extern "C" int g2048_gae_flat_v1(const float* d_rewards, const float* d_values, const uint8_t* d_dones, int64_t n,
                              double gamma, double lambda_gae, float* d_adv, float* d_ret, void* d_scan_state,
                              double* d_moments, void* stream) {
    G2048_REQUIRE(n >= 0, "gae_flat: n");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_rewards && d_values && d_dones && d_adv && d_ret && d_scan_state, "gae_flat: pointers");
    const int64_t n_tiles = (n + GAE_TILE - 1) / GAE_TILE;
    G2048_REQUIRE(n_tiles <= 0x7FFFFFFFll, "gae_flat: too many steps");
    static bool carveout_on[64] = {false};  // 14 CTAs of 15.5 KiB per SM need the shared-memory-heavy L1 split
    bool* carveout_set = device_once_flag(carveout_on);
    if (carveout_set && !*carveout_set) {
        cudaFuncSetAttribute(gae_flat_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        *carveout_set = true;
    }
    gae_flat_kernel<<<(unsigned)n_tiles, GAE_THREADS, 0, (cudaStream_t)stream>>>(
        d_rewards, d_values, d_dones, n, n_tiles, (float)gamma, (float)(gamma * lambda_gae), d_adv, d_ret,
        (GaeScratch*)d_scan_state, d_moments);
    G2048_CHECK_LAUNCH("gae_flat");
    return G2048_OK;
}
#endif  // G2048_LEGACY_KERNELS

extern "C" int g2048_gae_time_major(const float* d_rewards, const float* d_values, const uint8_t* d_rec_meta,
                                    int64_t t_steps, int64_t n, const float* d_bootstrap, double gamma,
                                    double lambda_gae, float* d_adv, float* d_ret, double* d_moments, void* stream) {
    G2048_REQUIRE(t_steps >= 0 && n >= 0, "gae_time_major: shape");
    if (t_steps == 0 || n == 0) return G2048_OK;
    G2048_REQUIRE(d_rewards && d_values && d_rec_meta && d_adv && d_ret, "gae_time_major: pointers");
    // G2048_GAE_TM_SPLIT=1 selects the form whose loads are spread over four warps (see the kernel): measured SLOWER
    // than one lane per env (C3 46 vs 34 us, 4 x C3 150 vs 113 us), so it is off unless asked for
    static const bool split = [] { const char* e = getenv("G2048_GAE_TM_SPLIT"); return e && atoi(e) != 0; }();
    if (split) {
        const int smem = (int)sizeof(GaeTsStage) * (t_steps > GAE_TS_CHUNK ? 2 : 1);
        static bool configured_on[64] = {false};
        bool* configured = device_once_flag(configured_on);
        if (!configured) return fail_arg("no CUDA device");
        if (!*configured) {
            const int rc = check_cuda(cudaFuncSetAttribute(gae_time_major_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                           2 * (int)sizeof(GaeTsStage)), "gae_time_major: shared memory attribute");
            if (rc) return rc;
            *configured = true;
        }
        gae_time_major_split_kernel<<<blocks_for(n, 32), 32 * GAE_TS_WARPS, smem, (cudaStream_t)stream>>>(
            d_rewards, d_values, d_rec_meta, t_steps, n, d_bootstrap, (float)gamma, (float)(gamma * lambda_gae), d_adv, d_ret,
            d_moments);
        G2048_CHECK_LAUNCH("gae_time_major");
        return G2048_OK;
    }
    // the bulk-copy ring form: whole CTAs of 128 envs, rows that the copy engine can address (16-byte aligned)
    static const bool ring = [] { const char* e = getenv("G2048_GAE_TM_RING"); return !(e && e[0] == '0'); }();
    const uintptr_t align_bits = (uintptr_t)d_rewards | (uintptr_t)d_values | (uintptr_t)d_rec_meta;
    if (ring && n % GAE_RING_ENVS == 0 && (align_bits & 15u) == 0) {
        static bool ring_configured_on[64] = {false};
        bool* configured = device_once_flag(ring_configured_on);
        if (!configured) return fail_arg("no CUDA device");
        if (!*configured) {
            int rc = check_cuda(cudaFuncSetAttribute(gae_time_major_ring_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                     cudaSharedmemCarveoutMaxShared), "gae_time_major: carveout");
            if (!rc) rc = check_cuda(cudaFuncSetAttribute(gae_time_major_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                          GAE_RING_STAGES * (int)sizeof(GaeRingStage)), "gae_time_major: shared memory attribute");
            if (rc) return rc;
            *configured = true;
        }
        gae_time_major_ring_kernel<<<(unsigned)(n / GAE_RING_ENVS), GAE_RING_ENVS, GAE_RING_STAGES * sizeof(GaeRingStage), (cudaStream_t)stream>>>(
            d_rewards, d_values, d_rec_meta, t_steps, n, d_bootstrap, (float)gamma, (float)(gamma * lambda_gae), d_adv, d_ret,
            d_moments);
    } else
        gae_time_major_kernel<<<blocks_for(n, GAE_TM_THREADS), GAE_TM_THREADS, 0, (cudaStream_t)stream>>>(
            d_rewards, d_values, d_rec_meta, t_steps, n, d_bootstrap, (float)gamma, (float)(gamma * lambda_gae), d_adv,
            d_ret, d_moments);
    G2048_CHECK_LAUNCH("gae_time_major");
    return G2048_OK;
}

extern "C" int g2048_normalize(float* d_x, int64_t n, const double* d_moments, int which, void* stream) {
    G2048_REQUIRE(n >= 0 && (which == 1 || which == 3), "normalize: arguments");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_x && d_moments, "normalize: pointers");
    const int sms = sm_count();
    if (sms <= 0) return fail_arg("normalize: no device");
    int64_t g = (n / 4 + 255) / 256 + 1;
    const int64_t cap = (int64_t)sms * 8 * 4;
    if (g > cap) g = cap;
    normalize_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(d_x, n, d_moments, which);
    G2048_CHECK_LAUNCH("normalize");
    return G2048_OK;
}

// Host-buffer form.  The caller's arrays are ordinary pageable memory (numpy).  The first version allocated and freed
// seven device buffers per call and used plain cudaMemcpy (the driver stages pageable memory itself, synchronously):
// 18.8 ms for 1e6 steps, 102 ms for 3.1e7 around 3 ms of kernels.  Now the calling thread keeps a workspace -- one
// stream, a grow-only device buffer and a StagedCopier (g2048_hostcopy.cuh) -- and the copies are pipelined through its
// pinned staging buffers: 1.5 ms and 74 ms
// (tools/probes/gae_host_bench.py; in the large case the time goes to first-touch page faults of the caller's freshly
// allocated 248 MB of outputs -- copying the staging chunks on four threads changed nothing).
namespace {

struct GaeHostWorkspace {
    int device = -1;
    cudaStream_t stream = nullptr;
    uint8_t* dev = nullptr;
    size_t dev_bytes = 0;
    StagedCopier copier;
    // everything goes back when the thread moves to another device or calls g2048_release_host_workspace()
    void release() {
        if (device < 0) return;
        int cur = -1;
        const bool switched = cudaGetDevice(&cur) == cudaSuccess && cur != device && cudaSetDevice(device) == cudaSuccess;
        if (dev) cudaFree(dev);
        if (stream) cudaStreamDestroy(stream);
        copier.release();
        if (switched) cudaSetDevice(cur);
        cudaGetLastError();
        dev = nullptr;
        stream = nullptr;
        dev_bytes = 0;
        device = -1;
    }
};

static thread_local GaeHostWorkspace t_gae_host_ws;

}  // namespace

namespace g2048 {
void release_play_host_workspace();  // g2048_env.cu
}

// Frees what the calling thread's *_host entry points keep between calls (a stream, a grow-only device buffer and
// 2 x 16 MiB of pinned staging memory each).  A thread that used them should call this before it ends.
extern "C" int g2048_release_host_workspace(void) {
    t_gae_host_ws.release();
    g2048::release_play_host_workspace();
    return G2048_OK;
}

extern "C" int g2048_gae_host(const float* h_rewards, const float* h_values, const uint8_t* h_dones, int64_t n,
                              double gamma, double lambda_gae, int normalize, float* h_adv, float* h_ret) {
    G2048_REQUIRE(n >= 0, "gae_host: n");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(h_rewards && h_values && h_dones && h_adv && h_ret, "gae_host: pointers");
    GaeHostWorkspace& ws = t_gae_host_ws;
    int rc = G2048_OK, dev = 0;
#define TRY(expr, where) do { rc = check_cuda((expr), where); if (rc) return rc; } while (0)
    TRY(cudaGetDevice(&dev), "gae_host: device");
    if (ws.device != dev) {  // first call on this thread, or the thread switched device
        ws.release();
        TRY(cudaStreamCreateWithFlags(&ws.stream, cudaStreamNonBlocking), "gae_host: stream");
        if ((rc = ws.copier.init())) return rc;
        ws.device = dev;
    }
    const auto align256 = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t fbytes = align256((size_t)n * 4), sb = (size_t)g2048_gae_flat_scratch_bytes(n);
    // layout: moments | scratch | rewards | values | adv | ret | dones
    const size_t off_m = 0, off_s = 256, off_r = off_s + align256(sb), off_v = off_r + fbytes, off_a = off_v + fbytes,
                 off_t = off_a + fbytes, off_d = off_t + fbytes, total = off_d + align256((size_t)n);
    if (total > ws.dev_bytes) {
        if (ws.dev) cudaFree(ws.dev);
        ws.dev = nullptr;
        ws.dev_bytes = 0;
        TRY(cudaMalloc((void**)&ws.dev, total), "gae_host: malloc");
        ws.dev_bytes = total;
    }
    double* d_m = (double*)(ws.dev + off_m);
    void* d_s = ws.dev + off_s;
    float *d_r = (float*)(ws.dev + off_r), *d_v = (float*)(ws.dev + off_v), *d_a = (float*)(ws.dev + off_a),
          *d_t = (float*)(ws.dev + off_t);
    uint8_t* d_d = ws.dev + off_d;
    TRY(cudaMemsetAsync(ws.dev, 0, off_r, ws.stream), "gae_host: memset");  // moments + scratch
    if ((rc = ws.copier.h2d(d_r, h_rewards, (size_t)n * 4, ws.stream))) return rc;
    if ((rc = ws.copier.h2d(d_v, h_values, (size_t)n * 4, ws.stream))) return rc;
    if ((rc = ws.copier.h2d(d_d, h_dones, (size_t)n, ws.stream))) return rc;
    if ((rc = g2048_gae_flat(d_r, d_v, d_d, n, gamma, lambda_gae, d_a, d_t, d_s, d_m, ws.stream))) return rc;
    if (normalize) {
        if ((rc = g2048_normalize(d_a, n, d_m, 1, ws.stream))) return rc;
        if ((rc = g2048_normalize(d_t, n, d_m, 3, ws.stream))) return rc;
    }
    if ((rc = ws.copier.d2h(h_adv, d_a, (size_t)n * 4, ws.stream))) return rc;
    if ((rc = ws.copier.d2h(h_ret, d_t, (size_t)n * 4, ws.stream))) return rc;
    TRY(cudaStreamSynchronize(ws.stream), "gae_host: sync");
#undef TRY
    return G2048_OK;
}
