// Env kernels: init, step, built-in policies, the persistent play-to-termination kernel and the
// lock-step recorded rollout.  All state of an env (board, mask, keys, counters) lives in the
// registers of the one thread that owns it; HBM sees only records and per-episode results.
#include "g2048_board.cuh"
#include <cstdlib>

#include "g2048_common.cuh"
#include "g2048_hostcopy.cuh"
#include "g2048_env.cuh"

namespace g2048 {

thread_local char g_last_error[512] = "";

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
        cached[dev] = n;
    }
    return cached[dev];
}

// ------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------
__global__ void threefry_kernel(const uint2* __restrict__ keys, const uint2* __restrict__ ctrs, int64_t n,
                                uint2* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Key y = threefry2x32(Key{keys[i].x, keys[i].y}, ctrs[i].x, ctrs[i].y);
    out[i] = make_uint2(y.a, y.b);
}

// key, sub = split(key), n_sub times.  Sequential by nature; the two children of a split are
// independent blocks, so one thread has ILP 2.  ~0.3 us per split.
template <int MODE>
__global__ void chain_kernel(uint32_t* key_io, int64_t n_sub, uint2* __restrict__ subs) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    Key k{key_io[0], key_io[1]};
    for (int64_t i = 0; i < n_sub; ++i) {
        Key nk, sub;
        split2<MODE>(k, nk, sub);
        subs[i] = make_uint2(sub.a, sub.b);
        k = nk;
    }
    key_io[0] = k.a;
    key_io[1] = k.b;
}

template <int MODE>
__global__ void split_keys_kernel(const uint32_t* __restrict__ sub, uint32_t batch_global, uint32_t env_lo, int64_t n,
                                  uint2* __restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Key k = env_key<MODE>(sub, batch_global, env_lo, i);
    keys[i] = make_uint2(k.a, k.b);
}

template <int MODE>
__global__ void env_init_kernel(const uint32_t* __restrict__ sub, uint32_t batch_global, uint32_t env_lo, int64_t n,
                                u64* __restrict__ boards, uint8_t* __restrict__ status) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Key k = env_key<MODE>(sub, batch_global, env_lo, i);
    const EnvState s = env_init<MODE>(k);
    boards[i] = s.board;
    status[i] = (uint8_t)s.status;
}

template <int MODE>
__global__ void env_step_kernel(u64* __restrict__ boards, uint8_t* __restrict__ status,
                                const int32_t* __restrict__ actions, const uint32_t* __restrict__ sub,
                                uint32_t batch_global, uint32_t env_lo, int64_t n, float* __restrict__ rewards) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    EnvState s{boards[i], status[i]};
    const Key k = env_key<MODE>(sub, batch_global, env_lo, i);
    const float r = env_step<MODE>(s, actions[i] & 3, k);
    boards[i] = s.board;
    status[i] = (uint8_t)s.status;
    if (rewards) rewards[i] = r;
}

__global__ void env_step_draws_kernel(u64* __restrict__ boards, uint8_t* __restrict__ status,
                                      const int32_t* __restrict__ actions, const uint32_t* __restrict__ bits_pos,
                                      const uint32_t* __restrict__ bits_val, int64_t n, float* __restrict__ rewards) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    EnvState s{boards[i], status[i]};
    const float r = env_step_draws(s, actions[i] & 3, bits_pos[i], bits_val[i]);
    boards[i] = s.board;
    status[i] = (uint8_t)s.status;
    if (rewards) rewards[i] = r;
}

template <int MODE, int POLICY>
__global__ void act_kernel(const uint8_t* __restrict__ status, const uint32_t* __restrict__ sub, uint32_t batch_global,
                           uint32_t env_lo, int64_t n, int32_t* __restrict__ actions, float* __restrict__ log_probs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t lm = status[i] & G2048_STATUS_MASK;
    int a;
    if (POLICY == G2048_POLICY_RANDOM) {
        const Key k = env_key<MODE>(sub, batch_global, env_lo, i);
        a = act_random<MODE>(k, lm);
        if (log_probs) log_probs[i] = act_random_log_prob(lm);
    } else {
        a = act_drul(lm);
    }
    actions[i] = a;
}

#ifdef G2048_LEGACY_KERNELS  // first-generation play kernel: built only into the tests' libg2048_legacy.so
// ------------------------------------------------------------------------------------------------
// persistent play-to-termination kernel
// ------------------------------------------------------------------------------------------------
// Each lane owns one env at a time and keeps it in registers until it terminates.  Finished lanes
// park until REFILL_MIN lanes of the warp are free (or the warp has nothing else to do), then the
// warp claims that many envs from a global queue with one atomic and initialises them together,
// so the divergent init path is paid once per ~REFILL_MIN episodes instead of once per episode.
constexpr int PLAY_THREADS = 256;
constexpr int REFILL_MIN = 6;

template <int MODE, int POLICY>
__global__ void __launch_bounds__(PLAY_THREADS)
play_kernel(const uint2* __restrict__ subs, int64_t n_subs, uint32_t batch_global, uint32_t env_lo, uint32_t n,
            unsigned long long* __restrict__ work, u64* __restrict__ final_boards, uint32_t* __restrict__ lengths,
            uint32_t* __restrict__ scores, unsigned long long* __restrict__ stats) {
    __shared__ unsigned long long s_stats[G2048_PLAY_STATS_WORDS];
    for (int i = threadIdx.x; i < G2048_PLAY_STATS_WORDS; i += blockDim.x) s_stats[i] = 0ull;
    __syncthreads();

    const unsigned lane = threadIdx.x & 31u;
    const Key init_sub{subs[0].x, subs[0].y};
    const uint32_t max_steps = (uint32_t)((n_subs - 1) / 2);

    // per-lane env state
    EnvState s{0ull, 0u};
    uint32_t e = 0, t = 0, score = 0;
    bool active = false;
    bool exhausted = false;  // warp-uniform: the queue has no more envs

    // per-lane statistics, flushed once at the end
    uint32_t st_episodes = 0, st_cut = 0, st_ovf = 0, st_longest = 0;
    unsigned long long st_steps = 0, st_score = 0, st_tile = 0, st_tile2 = 0;

    while (true) {
        const unsigned idle = __ballot_sync(0xFFFFFFFFu, !active);
        if (idle) {
            const int n_idle = __popc(idle);
            if (!exhausted && (n_idle >= REFILL_MIN || idle == 0xFFFFFFFFu)) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(work, (unsigned long long)n_idle);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (base + (unsigned long long)n_idle >= (unsigned long long)n) exhausted = true;
                if (!active) {
                    const unsigned long long mine = base + (unsigned long long)__popc(idle & ((1u << lane) - 1u));
                    if (mine < (unsigned long long)n) {
                        e = (uint32_t)mine;
                        s = env_init<MODE>(split_at<MODE>(init_sub, batch_global, env_lo + e));
                        t = 0;
                        score = 0;
                        active = true;
                    }
                }
            }
            if (__ballot_sync(0xFFFFFFFFu, active) == 0u) break;  // exhausted and nothing running
        }
        if (active) {
            const uint2 sa = __ldg(&subs[1 + 2 * (int64_t)t]);
            const uint2 ss = __ldg(&subs[2 + 2 * (int64_t)t]);
            const uint32_t lm = s.status & G2048_STATUS_MASK;
            int a;
            if (POLICY == G2048_POLICY_RANDOM) {
                a = act_random<MODE>(split_at<MODE>(Key{sa.x, sa.y}, batch_global, env_lo + e), lm);
            } else {
                a = act_drul(lm);
            }
            const float r = env_step<MODE>(s, a, split_at<MODE>(Key{ss.x, ss.y}, batch_global, env_lo + e));
            score += (r > 0.0f) ? (uint32_t)r : 0u;
            ++t;
            const bool done = (s.status & G2048_STATUS_DONE) != 0u;
            const bool cut = !done && t >= max_steps;
            if (done || cut) {
                if (final_boards) final_boards[e] = s.board;
                if (lengths) lengths[e] = t;
                if (scores) scores[e] = score;
                const uint32_t me = max_exponent(s.board);
                const unsigned long long tile = 1ull << me;
                st_episodes += 1;
                st_steps += t;
                st_score += score;
                st_cut += cut ? 1u : 0u;
                st_ovf += (s.status & G2048_STATUS_OVERFLOW) ? 1u : 0u;
                st_longest = max(st_longest, t);
                st_tile += tile;
                st_tile2 += tile * tile;
                atomicAdd(&s_stats[16 + me], 1ull);
                active = false;
            }
        }
    }

    atomicAdd(&s_stats[0], (unsigned long long)st_episodes);
    atomicAdd(&s_stats[1], st_steps);
    atomicAdd(&s_stats[2], st_score);
    atomicAdd(&s_stats[3], (unsigned long long)st_cut);
    atomicAdd(&s_stats[4], (unsigned long long)st_ovf);
    atomicMax(&s_stats[5], (unsigned long long)st_longest);
    atomicAdd(&s_stats[6], st_tile);
    atomicAdd(&s_stats[7], st_tile2);
    __syncthreads();
    for (int i = threadIdx.x; i < G2048_PLAY_STATS_WORDS; i += blockDim.x) {
        const unsigned long long v = s_stats[i];
        if (v) {
            if (i == 5) atomicMax(&stats[i], v);
            else atomicAdd(&stats[i], v);
        }
    }
}
#endif  // G2048_LEGACY_KERNELS

// ------------------------------------------------------------------------------------------------
// lock-step recorded rollout (trajectory mode)
// ------------------------------------------------------------------------------------------------
template <int MODE, int POLICY>
__global__ void __launch_bounds__(256)
rollout_steps_kernel(u64* __restrict__ boards, uint8_t* __restrict__ status, const uint2* __restrict__ subs,
                     int n_steps, uint32_t t0, uint32_t batch_global, uint32_t env_lo, int64_t n,
                     u64* __restrict__ rec_boards, uint8_t* __restrict__ rec_meta, float* __restrict__ rec_rewards,
                     float* __restrict__ rec_log_probs, unsigned long long* __restrict__ counters,
                     const int64_t* __restrict__ env_ids, int64_t n_rows) {
    // env_ids (live form): thread `row` serves env env_ids[row]; an env that finishes leaves the loop, and the record
    // slots of the steps after its end are not written (the reference's records repeat the frozen state there).
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long live_steps = 0, reward_sum = 0;
    if (row < n_rows) {  // no early return: every lane of the warp takes part in the shuffles below
        const int64_t i = env_ids ? env_ids[row] : row;
        EnvState s{boards[i], status[i]};
        const uint32_t e = env_lo + (uint32_t)i;
        for (int t = 0; t < n_steps; ++t) {
            const uint2 sa = __ldg(&subs[2 * t]);
            const uint2 ss = __ldg(&subs[2 * t + 1]);
            const uint32_t lm = s.status & G2048_STATUS_MASK;
            const bool was_done = (s.status & G2048_STATUS_DONE) != 0u;
            int a;
            float lp = 0.0f;
            if (POLICY == G2048_POLICY_RANDOM) {
                a = act_random<MODE>(split_at<MODE>(Key{sa.x, sa.y}, batch_global, e), lm);
                lp = act_random_log_prob(lm);
            } else {
                a = act_drul(lm);
            }
            const u64 pre = s.board;
            const float r = env_step<MODE>(s, a, split_at<MODE>(Key{ss.x, ss.y}, batch_global, e));
            const bool done = (s.status & G2048_STATUS_DONE) != 0u;
            const int64_t o = (int64_t)t * n + i;
            rec_boards[o] = pre;
            rec_meta[o] = (uint8_t)((uint32_t)a | (lm << 2) | (done ? 0x40u : 0u));
            rec_rewards[o] = r;
            if (rec_log_probs) rec_log_probs[o] = lp;
            if (!was_done) {
                live_steps += 1;
                if (r > 0.0f) reward_sum += (unsigned long long)r;
                if (done) {
                    atomicAdd(&counters[0], 1ull);
                    atomicMax(&counters[1], (unsigned long long)(t0 + (uint32_t)t + 1u));
                }
            }
            if (env_ids && done) break;
        }
        boards[i] = s.board;
        status[i] = (uint8_t)s.status;
    }
    // warp-aggregate the two sums
    for (int off = 16; off > 0; off >>= 1) {
        live_steps += __shfl_down_sync(0xFFFFFFFFu, live_steps, off);
        reward_sum += __shfl_down_sync(0xFFFFFFFFu, reward_sum, off);
    }
    if ((threadIdx.x & 31) == 0) {
        if (live_steps) atomicAdd(&counters[2], live_steps);
        if (reward_sum) atomicAdd(&counters[3], reward_sum);
    }
}

// ------------------------------------------------------------------------------------------------
// replay of a few envs of a batch (the state after k steps of the built-in policies)
// ------------------------------------------------------------------------------------------------
template <int MODE, int POLICY>
__global__ void __launch_bounds__(128)
replay_envs_kernel(const uint2* __restrict__ subs, uint32_t batch_global, const int64_t* __restrict__ env_ids,
                   const uint32_t* __restrict__ steps, int64_t m, u64* __restrict__ boards, uint8_t* __restrict__ status) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint32_t e = (uint32_t)env_ids[i];
    EnvState s = env_init<MODE>(split_at<MODE>(Key{subs[0].x, subs[0].y}, batch_global, e));
    const uint32_t k = steps[i];
    for (uint32_t t = 0; t < k; ++t) {
        const uint2 sa = __ldg(&subs[1 + 2 * (int64_t)t]);
        const uint2 ss = __ldg(&subs[2 + 2 * (int64_t)t]);
        const uint32_t lm = s.status & G2048_STATUS_MASK;
        const int a = (POLICY == G2048_POLICY_RANDOM) ? act_random<MODE>(split_at<MODE>(Key{sa.x, sa.y}, batch_global, e), lm)
                                                      : act_drul(lm);
        env_step<MODE>(s, a, split_at<MODE>(Key{ss.x, ss.y}, batch_global, e));
    }
    boards[i] = s.board;
    if (status) status[i] = (uint8_t)s.status;
}

// ------------------------------------------------------------------------------------------------
// integer-issue probe
// ------------------------------------------------------------------------------------------------
__global__ void int_peak_kernel(int iters, uint32_t* __restrict__ sink) {
    uint32_t a0 = threadIdx.x, b0 = blockIdx.x + 1u, a1 = a0 ^ 0x9E3779B9u, b1 = b0 + 0x7F4A7C15u;
    uint32_t a2 = a0 + 17u, b2 = b0 ^ 0x85EBCA6Bu, a3 = a1 + 29u, b3 = b1 ^ 0xC2B2AE35u;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#define G2048_MIX(a, b, r) a += b; b = rotl32(b, r); b ^= a;
        G2048_MIX(a0, b0, 13) G2048_MIX(a1, b1, 13) G2048_MIX(a2, b2, 13) G2048_MIX(a3, b3, 13)
        G2048_MIX(a0, b0, 15) G2048_MIX(a1, b1, 15) G2048_MIX(a2, b2, 15) G2048_MIX(a3, b3, 15)
        G2048_MIX(a0, b0, 26) G2048_MIX(a1, b1, 26) G2048_MIX(a2, b2, 26) G2048_MIX(a3, b3, 26)
        G2048_MIX(a0, b0, 6) G2048_MIX(a1, b1, 6) G2048_MIX(a2, b2, 6) G2048_MIX(a3, b3, 6)
#undef G2048_MIX
    }
    sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ b0 ^ a1 ^ b1 ^ a2 ^ b2 ^ a3 ^ b3;
}

}  // namespace g2048

// =================================================================================================
// C ABI
// =================================================================================================
using namespace g2048;

#define DISPATCH_MODE(rng_mode, CALL)                                   \
    if ((rng_mode) == G2048_RNG_PARTITIONABLE) { CALL(G2048_RNG_PARTITIONABLE); } \
    else { CALL(G2048_RNG_ORIGINAL); }

static inline bool valid_mode(int m) { return m == G2048_RNG_ORIGINAL || m == G2048_RNG_PARTITIONABLE; }
static inline bool valid_batch(int64_t batch_global, int64_t env_lo, int64_t n) {
    return batch_global > 0 && batch_global <= 0x7FFFFFFFll && env_lo >= 0 && n >= 0 && env_lo + n <= batch_global;
}
// batch_global == 0: d_sub is an explicit (n,2) key array
static inline bool valid_batch_or_keys(int64_t batch_global, int64_t env_lo, int64_t n) {
    return (batch_global == 0 && env_lo == 0 && n >= 0) || valid_batch(batch_global, env_lo, n);
}

extern "C" int g2048_version(void) { return 100; }
extern "C" const char* g2048_last_error(void) { return g_last_error; }
extern "C" int g2048_device_sm_count(void) { return sm_count(); }

extern "C" int g2048_threefry2x32(const uint32_t* d_keys, const uint32_t* d_ctrs, int64_t n, uint32_t* d_out,
                                  void* stream) {
    G2048_REQUIRE(n >= 0 && (n == 0 || (d_keys && d_ctrs && d_out)), "threefry2x32");
    if (n == 0) return G2048_OK;
    threefry_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const uint2*)d_keys, (const uint2*)d_ctrs, n,
                                                                         (uint2*)d_out);
    G2048_CHECK_LAUNCH("threefry2x32");
    return G2048_OK;
}

extern "C" int g2048_chain_advance(uint32_t* d_key_io, int rng_mode, int64_t n_sub, uint32_t* d_subs, void* stream) {
    G2048_REQUIRE(d_key_io && valid_mode(rng_mode) && n_sub >= 0 && (n_sub == 0 || d_subs), "chain_advance");
    if (n_sub == 0) return G2048_OK;
#define CALL(M) chain_kernel<M><<<1, 32, 0, (cudaStream_t)stream>>>(d_key_io, n_sub, (uint2*)d_subs)
    DISPATCH_MODE(rng_mode, CALL)
#undef CALL
    G2048_CHECK_LAUNCH("chain_advance");
    return G2048_OK;
}

extern "C" int g2048_split_keys(const uint32_t* d_sub, int64_t batch_global, int64_t env_lo, int64_t n, int rng_mode,
                                uint32_t* d_keys, void* stream) {
    G2048_REQUIRE(d_sub && valid_mode(rng_mode) && valid_batch_or_keys(batch_global, env_lo, n), "split_keys");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_keys, "split_keys: d_keys");
#define CALL(M)                                                                                         \
    split_keys_kernel<M><<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(d_sub, (uint32_t)batch_global, \
                                                                              (uint32_t)env_lo, n, (uint2*)d_keys)
    DISPATCH_MODE(rng_mode, CALL)
#undef CALL
    G2048_CHECK_LAUNCH("split_keys");
    return G2048_OK;
}

extern "C" int g2048_env_init(const uint32_t* d_sub, int64_t batch_global, int64_t env_lo, int64_t n, int rng_mode,
                              uint64_t* d_boards, uint8_t* d_status, void* stream) {
    G2048_REQUIRE(d_sub && valid_mode(rng_mode) && valid_batch_or_keys(batch_global, env_lo, n), "env_init");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_boards && d_status, "env_init: outputs");
#define CALL(M)                                                                                        \
    env_init_kernel<M><<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(d_sub, (uint32_t)batch_global,  \
                                                                            (uint32_t)env_lo, n, (u64*)d_boards, d_status)
    DISPATCH_MODE(rng_mode, CALL)
#undef CALL
    G2048_CHECK_LAUNCH("env_init");
    return G2048_OK;
}

extern "C" int g2048_env_step(uint64_t* d_boards, uint8_t* d_status, const int32_t* d_actions, const uint32_t* d_sub,
                              int64_t batch_global, int64_t env_lo, int64_t n, int rng_mode, float* d_rewards,
                              void* stream) {
    G2048_REQUIRE(d_sub && valid_mode(rng_mode) && valid_batch_or_keys(batch_global, env_lo, n), "env_step");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_boards && d_status && d_actions, "env_step: state/actions");
#define CALL(M)                                                                                         \
    env_step_kernel<M><<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(                            \
        (u64*)d_boards, d_status, d_actions, d_sub, (uint32_t)batch_global, (uint32_t)env_lo, n, d_rewards)
    DISPATCH_MODE(rng_mode, CALL)
#undef CALL
    G2048_CHECK_LAUNCH("env_step");
    return G2048_OK;
}

extern "C" int g2048_env_step_draws(uint64_t* d_boards, uint8_t* d_status, const int32_t* d_actions,
                                    const uint32_t* d_bits_pos, const uint32_t* d_bits_val, int64_t n,
                                    float* d_rewards, void* stream) {
    G2048_REQUIRE(n >= 0, "env_step_draws: n");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_boards && d_status && d_actions && d_bits_pos && d_bits_val, "env_step_draws: pointers");
    env_step_draws_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>((u64*)d_boards, d_status, d_actions,
                                                                               d_bits_pos, d_bits_val, n, d_rewards);
    G2048_CHECK_LAUNCH("env_step_draws");
    return G2048_OK;
}

extern "C" int g2048_act(int policy, const uint8_t* d_status, const uint32_t* d_sub, int64_t batch_global,
                         int64_t env_lo, int64_t n, int rng_mode, int32_t* d_actions, float* d_log_probs,
                         void* stream) {
    G2048_REQUIRE(policy == G2048_POLICY_RANDOM || policy == G2048_POLICY_DRUL, "act: policy");
    G2048_REQUIRE(valid_mode(rng_mode) && valid_batch_or_keys(batch_global, env_lo, n), "act");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_status && d_actions && (policy == G2048_POLICY_DRUL || d_sub), "act: pointers");
    const unsigned g = blocks_for(n, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (policy == G2048_POLICY_RANDOM) {
#define CALL(M) act_kernel<M, G2048_POLICY_RANDOM><<<g, 256, 0, st>>>(d_status, d_sub, (uint32_t)batch_global, (uint32_t)env_lo, n, d_actions, d_log_probs)
        DISPATCH_MODE(rng_mode, CALL)
#undef CALL
    } else {
        act_kernel<G2048_RNG_PARTITIONABLE, G2048_POLICY_DRUL><<<g, 256, 0, st>>>(d_status, d_sub, (uint32_t)batch_global,
                                                                                 (uint32_t)env_lo, n, d_actions, nullptr);
    }
    G2048_CHECK_LAUNCH("act");
    return G2048_OK;
}

#ifdef G2048_LEGACY_KERNELS
template <int MODE, int POLICY>
static int launch_play(const uint32_t* d_subs, int64_t n_subs, int64_t batch_global, int64_t env_lo, int64_t n,
                       uint64_t* d_work, uint64_t* d_final_boards, uint32_t* d_lengths, uint32_t* d_scores,
                       uint64_t* d_stats, cudaStream_t st) {
    int per_sm = 0;
    int rc = check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, play_kernel<MODE, POLICY>, PLAY_THREADS, 0),
                        "play: occupancy");
    if (rc) return rc;
    const int sms = sm_count();
    if (sms <= 0 || per_sm <= 0) return fail_arg("play: no device");
    int64_t grid = (int64_t)sms * per_sm;
    const int64_t needed = (n + PLAY_THREADS - 1) / PLAY_THREADS;
    if (grid > needed) grid = needed;
    play_kernel<MODE, POLICY><<<(unsigned)grid, PLAY_THREADS, 0, st>>>(
        (const uint2*)d_subs, n_subs, (uint32_t)batch_global, (uint32_t)env_lo, (uint32_t)n,
        (unsigned long long*)d_work, (u64*)d_final_boards, d_lengths, d_scores, (unsigned long long*)d_stats);
    return check_cuda(cudaGetLastError(), "play");
}

// First-generation play kernel (park-and-refill, per-step reward loop).  Kept as a measured baseline
// for the current kernel in g2048_play.cu; same arguments and results as g2048_play.
extern "C" int g2048_play_v1(int policy, const uint32_t* d_subs, int64_t n_subs, int64_t batch_global, int64_t env_lo,
                             int64_t n, int rng_mode, uint64_t* d_work, uint64_t* d_final_boards, uint32_t* d_lengths,
                             uint32_t* d_scores, uint64_t* d_stats, void* stream) {
    G2048_REQUIRE(policy == G2048_POLICY_RANDOM || policy == G2048_POLICY_DRUL, "play: policy");
    G2048_REQUIRE(valid_mode(rng_mode) && valid_batch(batch_global, env_lo, n), "play: batch");
    G2048_REQUIRE(n_subs >= 3 && d_subs && d_work && d_stats, "play: pointers");
    if (n == 0) return G2048_OK;
    cudaStream_t st = (cudaStream_t)stream;
#define ARGS d_subs, n_subs, batch_global, env_lo, n, d_work, d_final_boards, d_lengths, d_scores, d_stats, st
    if (policy == G2048_POLICY_RANDOM) {
        if (rng_mode == G2048_RNG_PARTITIONABLE) return launch_play<G2048_RNG_PARTITIONABLE, G2048_POLICY_RANDOM>(ARGS);
        return launch_play<G2048_RNG_ORIGINAL, G2048_POLICY_RANDOM>(ARGS);
    }
    if (rng_mode == G2048_RNG_PARTITIONABLE) return launch_play<G2048_RNG_PARTITIONABLE, G2048_POLICY_DRUL>(ARGS);
    return launch_play<G2048_RNG_ORIGINAL, G2048_POLICY_DRUL>(ARGS);
#undef ARGS
}
#endif  // G2048_LEGACY_KERNELS

// device-side address of a host pointer in pinned, mapped memory; nullptr for pageable memory
static void* mapped_device_pointer(const void* host) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, host) != cudaSuccess) {
        cudaGetLastError();  // older drivers report pageable memory as an error: clear it
        return nullptr;
    }
    return attr.type == cudaMemoryTypeHost ? attr.devicePointer : nullptr;
}

// Workspace of the calling thread for the *_host play entry points.
struct PlayHostWorkspace {
    int device = -1;
    cudaStream_t stream = nullptr;
    uint8_t* buf = nullptr;
    size_t bytes = 0;
    StagedCopier copier;  // result arrays in pageable memory (numpy) go through pinned staging buffers
    // everything goes back when the thread moves to another device or calls g2048_release_host_workspace()
    void release() {
        if (device < 0) return;
        int cur = -1;
        const bool switched = cudaGetDevice(&cur) == cudaSuccess && cur != device && cudaSetDevice(device) == cudaSuccess;
        if (buf) cudaFree(buf);
        if (stream) cudaStreamDestroy(stream);
        copier.release();
        if (switched) cudaSetDevice(cur);
        cudaGetLastError();
        buf = nullptr;
        stream = nullptr;
        bytes = 0;
        device = -1;
    }
};
static thread_local PlayHostWorkspace t_play_host_ws;

namespace g2048 {
void release_play_host_workspace() { t_play_host_ws.release(); }
}

static int play_host_impl(int policy, uint64_t seed, uint32_t* h_key_io, int64_t batch_global, int64_t env_lo,
                          int64_t n, int rng_mode, uint64_t* h_final_boards, uint32_t* h_lengths,
                          uint32_t* h_scores, G2048EpisodeResult* h_results, uint64_t* h_stats) {
    G2048_REQUIRE(policy == G2048_POLICY_RANDOM || policy == G2048_POLICY_DRUL, "play_host: policy");
    G2048_REQUIRE(valid_mode(rng_mode) && valid_batch(batch_global, env_lo, n), "play_host: batch");
    uint32_t key[2] = {(uint32_t)(seed >> 32), (uint32_t)(seed & 0xFFFFFFFFull)};
    if (h_key_io) { key[0] = h_key_io[0]; key[1] = h_key_io[1]; }

    // Workspace of the calling thread on the current device: one stream and one grow-only device
    // buffer, kept across calls so that a call costs copies and kernels, not cudaMalloc/cudaFree.
    PlayHostWorkspace& ws = t_play_host_ws;
    static const bool zero_copy = [] { const char* e = getenv("G2048_PLAY_HOST_ZEROCOPY"); return !(e && e[0] == '0'); }();
    int rc = G2048_OK;
    int dev = 0;
    uint64_t stats[G2048_PLAY_STATS_WORDS];
    // Loop steps of keys to generate: the chain kernel is serial (0.14 us per split, 284 us for 1 024 steps in front of
    // a 9 ms play kernel), so after the first call the length follows the longest episode this thread has seen with
    // this policy (+25 %, rounded up to 128) instead of a constant 1 024; a batch that outlives its keys is replayed
    // with four times as many, as before.  Results do not depend on the length.
    static thread_local int64_t longest_seen[2] = {0, 0};
    int64_t max_steps = 1024;
    if (longest_seen[policy] > 0) {
        max_steps = ((longest_seen[policy] * 5 / 4 + 127) / 128) * 128;
        if (max_steps < 256) max_steps = 256;
    }
#define TRY(expr, where) do { rc = check_cuda((expr), where); if (rc) return rc; } while (0)
    TRY(cudaGetDevice(&dev), "play_host: device");
    if (ws.device != dev) {  // first call on this thread, or the thread switched device
        ws.release();
        TRY(cudaStreamCreateWithFlags(&ws.stream, cudaStreamNonBlocking), "play_host: stream");
        if ((rc = ws.copier.init())) return rc;
        ws.device = dev;
    }
    cudaStream_t st = ws.stream;
    const auto align256 = [](size_t x) { return (x + 255) & ~(size_t)255; };
    while (true) {
        const int64_t n_subs = 1 + 2 * max_steps;
        // layout: key | work | stats | subs | boards | lengths | scores
        const size_t off_key = 0, off_work = 256, off_stats = 512, off_subs = 1024;
        const size_t off_boards = off_subs + align256((size_t)n_subs * 8);
        const size_t off_len = off_boards + align256(h_final_boards ? (size_t)n * 8 : 0);
        const size_t off_score = off_len + align256(h_lengths ? (size_t)n * 4 : 0);
        const size_t off_res = off_score + align256(h_scores ? (size_t)n * 4 : 0);
        const size_t total = off_res + align256(h_results ? (size_t)n * sizeof(G2048EpisodeResult) : 0);
        if (total > ws.bytes) {
            if (ws.buf) cudaFree(ws.buf);
            ws.buf = nullptr;
            ws.bytes = 0;
            TRY(cudaMalloc(&ws.buf, total), "play_host: malloc");
            ws.bytes = total;
        }
        uint32_t* d_key = (uint32_t*)(ws.buf + off_key);
        uint64_t* d_work = (uint64_t*)(ws.buf + off_work);
        uint64_t* d_stats = (uint64_t*)(ws.buf + off_stats);
        uint32_t* d_subs = (uint32_t*)(ws.buf + off_subs);
        uint64_t* d_boards = h_final_boards ? (uint64_t*)(ws.buf + off_boards) : nullptr;
        uint32_t* d_len = h_lengths ? (uint32_t*)(ws.buf + off_len) : nullptr;
        uint32_t* d_score = h_scores ? (uint32_t*)(ws.buf + off_score) : nullptr;
        G2048EpisodeResult* d_res = h_results ? (G2048EpisodeResult*)(ws.buf + off_res) : nullptr;
        // A result array in pinned (page-locked, mapped) host memory is written by the kernel itself: an episode's
        // 16 bytes of results leave over PCIe when the episode ends, overlapped with the rest of the batch, instead
        // of 32 MiB of copies after the kernel (G2048_PLAY_HOST_ZEROCOPY=0 restores the copies).
        bool zc_boards = false, zc_len = false, zc_score = false, zc_res = false;
        if (zero_copy) {
            void* m = nullptr;
            if (d_res && (m = mapped_device_pointer(h_results)) && ((uintptr_t)m & 15u) == 0) { d_res = (G2048EpisodeResult*)m; zc_res = true; }
            if (d_boards && (m = mapped_device_pointer(h_final_boards))) { d_boards = (uint64_t*)m; zc_boards = true; }
            if (d_len && (m = mapped_device_pointer(h_lengths))) { d_len = (uint32_t*)m; zc_len = true; }
            if (d_score && (m = mapped_device_pointer(h_scores))) { d_score = (uint32_t*)m; zc_score = true; }
        }
        TRY(cudaMemcpyAsync(d_key, key, sizeof(key), cudaMemcpyHostToDevice, st), "play_host: h2d key");
        TRY(cudaMemsetAsync(d_work, 0, 768, st), "play_host: memset");  // work + stats
        rc = g2048_chain_advance(d_key, rng_mode, n_subs, d_subs, st);
        if (rc) return rc;
        if (h_results)  // one 16-byte record per env: a third of the stores (and of the PCIe packets when they go to the host)
            rc = g2048_play_packed(policy, d_subs, n_subs, batch_global, env_lo, n, rng_mode, d_work, d_res, d_stats, st);
        else
            rc = g2048_play(policy, d_subs, n_subs, batch_global, env_lo, n, rng_mode, d_work, d_boards, d_len, d_score,
                            d_stats, st);
        if (rc) return rc;
        // results are copied optimistically; a batch that outlived its keys is replayed below
        TRY(cudaMemcpyAsync(stats, d_stats, sizeof(stats), cudaMemcpyDeviceToHost, st), "play_host: d2h stats");
        if (d_boards && !zc_boards && n && (rc = ws.copier.d2h(h_final_boards, d_boards, n * sizeof(uint64_t), st))) return rc;
        if (d_len && !zc_len && n && (rc = ws.copier.d2h(h_lengths, d_len, n * sizeof(uint32_t), st))) return rc;
        if (d_score && !zc_score && n && (rc = ws.copier.d2h(h_scores, d_score, n * sizeof(uint32_t), st))) return rc;
        if (d_res && !zc_res && n && (rc = ws.copier.d2h(h_results, d_res, n * sizeof(G2048EpisodeResult), st))) return rc;
        TRY(cudaStreamSynchronize(st), "play_host: sync");
        if (stats[3] == 0 || max_steps >= (1 << 20)) break;
        max_steps *= 4;  // some episode outlived the chain: replay with a longer one
    }
    if (stats[3] == 0 && (int64_t)stats[5] > longest_seen[policy]) longest_seen[policy] = (int64_t)stats[5];
    if (h_stats) for (int i = 0; i < G2048_PLAY_STATS_WORDS; ++i) h_stats[i] = stats[i];
    if (h_key_io) {
        // the reference's runner holds the chain key after 1 + 2*T splits, T = longest episode
        uint32_t* d_key = (uint32_t*)ws.buf;
        uint32_t* d_subs = (uint32_t*)(ws.buf + 1024);
        const int64_t used = 1 + 2 * (int64_t)stats[5];
        TRY(cudaMemcpyAsync(d_key, key, sizeof(key), cudaMemcpyHostToDevice, st), "play_host: h2d key");
        rc = g2048_chain_advance(d_key, rng_mode, used, d_subs, st);
        if (rc) return rc;
        TRY(cudaMemcpyAsync(h_key_io, d_key, sizeof(key), cudaMemcpyDeviceToHost, st), "play_host: d2h key");
        TRY(cudaStreamSynchronize(st), "play_host: sync");
    }
#undef TRY
    return rc;
}

extern "C" int g2048_play_host(int policy, uint64_t seed, uint32_t* h_key_io, int64_t batch_global, int64_t env_lo,
                               int64_t n, int rng_mode, uint64_t* h_final_boards, uint32_t* h_lengths,
                               uint32_t* h_scores, uint64_t* h_stats) {
    return play_host_impl(policy, seed, h_key_io, batch_global, env_lo, n, rng_mode, h_final_boards, h_lengths, h_scores,
                          nullptr, h_stats);
}

extern "C" int g2048_play_host_packed(int policy, uint64_t seed, uint32_t* h_key_io, int64_t batch_global, int64_t env_lo,
                                      int64_t n, int rng_mode, G2048EpisodeResult* h_results, uint64_t* h_stats) {
    return play_host_impl(policy, seed, h_key_io, batch_global, env_lo, n, rng_mode, nullptr, nullptr, nullptr, h_results,
                          h_stats);
}

static int launch_rollout_steps(int policy, uint64_t* d_boards, uint8_t* d_status, const uint32_t* d_subs,
                                int64_t n_steps, int64_t t0, int64_t batch_global, int64_t env_lo, int64_t n,
                                int rng_mode, const int64_t* d_env_ids, int64_t n_live, uint64_t* d_rec_boards,
                                uint8_t* d_rec_meta, float* d_rec_rewards, float* d_rec_log_probs, uint64_t* d_counters,
                                void* stream) {
    G2048_REQUIRE(!d_env_ids || (n_live >= 0 && n_live <= n), "rollout_steps: live rows");
    G2048_REQUIRE(policy == G2048_POLICY_RANDOM || policy == G2048_POLICY_DRUL, "rollout_steps: policy");
    G2048_REQUIRE(valid_mode(rng_mode) && valid_batch(batch_global, env_lo, n), "rollout_steps: batch");
    G2048_REQUIRE(n_steps >= 0 && n_steps <= 0x7FFFFFFF && t0 >= 0, "rollout_steps: steps");
    if (n == 0 || n_steps == 0) return G2048_OK;
    G2048_REQUIRE(d_boards && d_status && d_subs && d_rec_boards && d_rec_meta && d_rec_rewards && d_counters,
                  "rollout_steps: pointers");
    const int64_t n_rows = d_env_ids ? n_live : n;
    if (n_rows == 0) return G2048_OK;
    // one warp per CTA while there are fewer warps than schedulers on the GPU: a small batch is latency-bound, and 1 024
    // envs as four 256-thread CTAs put eight warps on each of four SMs (see launch_play2)
    const int sms_now = sm_count();
    const unsigned threads = (sms_now > 0 && n_rows <= (int64_t)sms_now * 4 * 32) ? 32u : 256u;
    const unsigned g = blocks_for(n_rows, threads);
    cudaStream_t st = (cudaStream_t)stream;
#define CALL_P(M, P)                                                                                          \
    rollout_steps_kernel<M, P><<<g, threads, 0, st>>>((u64*)d_boards, d_status, (const uint2*)d_subs, (int)n_steps,    \
                                                  (uint32_t)t0, (uint32_t)batch_global, (uint32_t)env_lo, n,        \
                                                  (u64*)d_rec_boards, d_rec_meta, d_rec_rewards, d_rec_log_probs,   \
                                                  (unsigned long long*)d_counters, d_env_ids, n_rows)
    if (policy == G2048_POLICY_RANDOM) {
#define CALL(M) CALL_P(M, G2048_POLICY_RANDOM)
        DISPATCH_MODE(rng_mode, CALL)
#undef CALL
    } else {
#define CALL(M) CALL_P(M, G2048_POLICY_DRUL)
        DISPATCH_MODE(rng_mode, CALL)
#undef CALL
    }
#undef CALL_P
    G2048_CHECK_LAUNCH("rollout_steps");
    return G2048_OK;
}

extern "C" int g2048_rollout_steps(int policy, uint64_t* d_boards, uint8_t* d_status, const uint32_t* d_subs,
                                   int64_t n_steps, int64_t t0, int64_t batch_global, int64_t env_lo, int64_t n,
                                   int rng_mode, uint64_t* d_rec_boards, uint8_t* d_rec_meta, float* d_rec_rewards,
                                   float* d_rec_log_probs, uint64_t* d_counters, void* stream) {
    return launch_rollout_steps(policy, d_boards, d_status, d_subs, n_steps, t0, batch_global, env_lo, n, rng_mode, nullptr, 0,
                                d_rec_boards, d_rec_meta, d_rec_rewards, d_rec_log_probs, d_counters, stream);
}

extern "C" int g2048_rollout_steps_live(int policy, uint64_t* d_boards, uint8_t* d_status, const uint32_t* d_subs,
                                        int64_t n_steps, int64_t t0, int64_t batch_global, int64_t env_lo, int64_t n,
                                        int rng_mode, const int64_t* d_env_ids, int64_t n_live, uint64_t* d_rec_boards,
                                        uint8_t* d_rec_meta, float* d_rec_rewards, float* d_rec_log_probs,
                                        uint64_t* d_counters, void* stream) {
    G2048_REQUIRE(d_env_ids || n_live == 0, "rollout_steps_live: env ids");
    if (n_live == 0) return G2048_OK;
    return launch_rollout_steps(policy, d_boards, d_status, d_subs, n_steps, t0, batch_global, env_lo, n, rng_mode, d_env_ids,
                                n_live, d_rec_boards, d_rec_meta, d_rec_rewards, d_rec_log_probs, d_counters, stream);
}

extern "C" int g2048_int_peak_probe(int blocks, int threads, int iters, uint32_t* d_sink, void* stream) {
    G2048_REQUIRE(blocks > 0 && threads > 0 && threads <= 1024 && iters > 0 && d_sink, "int_peak_probe");
    int_peak_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, d_sink);
    G2048_CHECK_LAUNCH("int_peak_probe");
    return G2048_OK;
}

// State of the listed envs (GLOBAL indices, int64) of BatchRunner(seed).run_*(batch_global) after d_steps[i] loop steps of
// the built-in policy -- e.g. the board an env held BEFORE its last step, which is what the reference's
// run_actions_max_tile reads for the envs that live until the last loop step (src/runs/run_actions_max_tile.py:61-64).
extern "C" int g2048_replay_envs(int policy, const uint32_t* d_subs, int64_t n_subs, int64_t batch_global,
                                 const int64_t* d_env_ids, const uint32_t* d_steps, int64_t m, int rng_mode,
                                 uint64_t* d_boards, uint8_t* d_status, void* stream) {
    G2048_REQUIRE(policy == G2048_POLICY_RANDOM || policy == G2048_POLICY_DRUL, "replay_envs: policy");
    G2048_REQUIRE(valid_mode(rng_mode) && batch_global > 0 && batch_global <= 0x7FFFFFFFll && m >= 0 && n_subs >= 1, "replay_envs");
    if (m == 0) return G2048_OK;
    G2048_REQUIRE(d_subs && d_env_ids && d_steps && d_boards, "replay_envs: pointers");
    const unsigned g = blocks_for(m, 128);
    cudaStream_t st = (cudaStream_t)stream;
#define CALL_P(M, P) replay_envs_kernel<M, P><<<g, 128, 0, st>>>((const uint2*)d_subs, (uint32_t)batch_global, d_env_ids, d_steps, m, (u64*)d_boards, d_status)
    if (policy == G2048_POLICY_RANDOM) {
#define CALL(M) CALL_P(M, G2048_POLICY_RANDOM)
        DISPATCH_MODE(rng_mode, CALL)
#undef CALL
    } else {
#define CALL(M) CALL_P(M, G2048_POLICY_DRUL)
        DISPATCH_MODE(rng_mode, CALL)
#undef CALL
    }
#undef CALL_P
    G2048_CHECK_LAUNCH("replay_envs");
    return G2048_OK;
}
