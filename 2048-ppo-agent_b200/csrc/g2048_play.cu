// Persistent play-to-termination kernel (second generation) -- the whole of
// src/runs/batch_runner.py:105-136 for act_randomly / act_drul plus the max-tile reduction of
// src/runs/run_actions_max_tile.py:61-69, for envs [env_lo, env_lo + n) of a global batch.
//
// What changed against g2048_play_v1 (ncu: ALU pipe 94.7 % busy, but only 23.2 of 32 lanes active):
//   * No lane ever waits.  A lane that finishes an episode takes the next env from the queue at the
//     top of the very next iteration, and the env's init (Pgx _init: two spawns on the empty board)
//     runs as two ordinary loop iterations ("phase 0/1") through the SAME instruction stream as a
//     game step, so initialising lanes do not diverge from playing lanes.  This works because the
//     Threefry blocks a random-policy step hashes anyway -- TF(k; (0,i)), i < 4, under the per-env
//     act key -- are exactly the blocks jax.random.split(k) is made of: an initialising lane feeds
//     the env's init key in place of the act key and reads its spawn key r_phase out of the same
//     blocks (both counter layouts).  The DRUL policy has no act-key work to share, so its
//     initialising lanes hash one extra block under a short divergent branch.
//   * No per-step reward.  The episode score is board_potential(final) - 4 * (#spawned 4-tiles),
//     which equals the sum of Pgx's merge rewards; the divergent per-merge loop is gone.
//   * Overflow (a 2^16 tile would be needed) is detected by looking for a 2^15 tile every 256 steps and
//     at the end of the episode: a second 2^15 tile cannot be built in fewer steps than that.
#include "g2048_board.cuh"
#include "g2048_common.cuh"
#include "g2048_env.cuh"
#include "g2048_play.cuh"

namespace g2048 {

constexpr int PLAY2_THREADS = 256;
// 0 restores round 1's launch for A/B timing: 256-thread CTAs at every batch size
#ifndef G2048_PLAY2_NARROW
#define G2048_PLAY2_NARROW 1
#endif

// The random draws of one loop iteration: the four bit words of the action draw (random policy) and the two of the
// spawn.  A pure function of (env, loop step, phase) -- the keys come from the launch's sub keys and the env index,
// never from the board.  (Making them one iteration AHEAD, next to the game logic they do not depend on, was built for
// launches in which a lane takes one env only: bit-exact, and worth 4 % at one warp per SM with the random policy,
// nothing with DRUL -- the warp issues in order and the two chains sit in different basic blocks.  Not kept; the
// same idea in the table kernel, where lanes claim envs all the time, was 14 % slower: see play3_step.)
struct Play2Draws {
    uint32_t bits[4];
    uint32_t pos, val;
};

template <int MODE, int POLICY>
__device__ __forceinline__ Play2Draws play2_draws(const uint2* __restrict__ subs, uint2 init_sub, uint32_t batch_global,
                                                  uint32_t env, uint32_t t, uint32_t phase) {
    Play2Draws r;
    const bool playing = phase == PHASE_PLAY;
    const uint2 ss = __ldg(&subs[2 + 2 * (int64_t)t]);
    Key kstep;
    if (POLICY == G2048_POLICY_RANDOM) {
        const uint2 sa = __ldg(&subs[1 + 2 * (int64_t)t]);
        const Key head = playing ? Key{sa.x, sa.y} : Key{init_sub.x, init_sub.y};
        const KeyBlocks<MODE> kb = key_blocks<MODE>(split_at<MODE>(head, batch_global, env));
#pragma unroll
        for (int i = 0; i < 4; ++i) r.bits[i] = kb.bits[i];
        const Key kplay = split_at<MODE>(Key{ss.x, ss.y}, batch_global, env);
        const Key kinit = (phase == PHASE_INIT0) ? kb.child[0] : kb.child[1];
        kstep = playing ? kplay : kinit;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) r.bits[i] = 0u;
        const Key head = playing ? Key{ss.x, ss.y} : Key{init_sub.x, init_sub.y};
        kstep = split_at<MODE>(head, batch_global, env);
        if (!playing) {  // r_phase = split(init key)[phase]
            Key c0, c1;
            split2<MODE>(kstep, c0, c1);
            kstep = (phase == PHASE_INIT0) ? c0 : c1;
        }
    }
    Key k1, k2;
    split2<MODE>(kstep, k1, k2);
    r.pos = bits_scalar<MODE>(k1);
    r.val = bits_scalar<MODE>(k2);
    return r;
}

template <int MODE, int POLICY>
__global__ void __launch_bounds__(PLAY2_THREADS)
play2_kernel(const uint2* __restrict__ subs, int64_t n_subs, uint32_t batch_global, uint32_t env_lo, uint32_t n,
             unsigned long long* __restrict__ work, u64* __restrict__ final_boards, uint32_t* __restrict__ lengths,
             uint32_t* __restrict__ scores, unsigned long long* __restrict__ stats, uint4* __restrict__ results) {
    __shared__ unsigned long long s_stats[G2048_PLAY_STATS_WORDS];
    for (int i = threadIdx.x; i < G2048_PLAY_STATS_WORDS; i += blockDim.x) s_stats[i] = 0ull;
    __syncthreads();

    const unsigned lane = threadIdx.x & 31u;
    const uint2 init_sub = subs[0];
    const uint32_t max_steps = (uint32_t)((n_subs - 1) / 2);

    u64 board = 0ull;
    uint32_t lm = 0, e = 0, t = 0, fours = 0, phase = PHASE_NONE;
    bool seen15 = false;
    bool exhausted = false;  // warp-uniform

    uint32_t st_episodes = 0, st_cut = 0, st_ovf = 0, st_longest = 0;
    unsigned long long st_steps = 0, st_score = 0, st_tile = 0, st_tile2 = 0;

    while (true) {
        // ---- hand the next envs of the queue to the lanes that have none --------------------------
        const unsigned want = __ballot_sync(0xFFFFFFFFu, phase == PHASE_NONE);
        if (want) {
            if (!exhausted) {
                const int cnt = __popc(want);
                unsigned long long base = 0;
                race_jitter();
                if (lane == 0) base = atomicAdd(work, (unsigned long long)cnt);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (base + (unsigned long long)cnt >= (unsigned long long)n) exhausted = true;
                if (phase == PHASE_NONE) {
                    const unsigned long long mine = base + (unsigned long long)__popc(want & ((1u << lane) - 1u));
                    if (mine < (unsigned long long)n) {
                        e = (uint32_t)mine;
                        board = 0ull;
                        lm = 0;
                        t = 0;
                        fours = 0;
                        seen15 = false;
                        phase = PHASE_INIT0;
                    }
                }
            }
            if (__ballot_sync(0xFFFFFFFFu, phase != PHASE_NONE) == 0u) break;
        }
        if (phase == PHASE_NONE) continue;  // only at the tail of the launch, when the queue is empty

        // ---- one iteration: a game step, or one of the two init spawns ------------------------------
        const bool playing = phase == PHASE_PLAY;
        const Play2Draws d = play2_draws<MODE, POLICY>(subs, init_sub, batch_global, env_lo + e, t, phase);
        const int action = (POLICY == G2048_POLICY_RANDOM) ? argmax_bits_legal(d.bits, lm) : act_drul(lm);
        const uint32_t bits_pos = d.pos, bits_val = d.val;

        uint32_t unused_reward = 0;
        bool unused_ovf = false;
        const u64 moved = move_board<false>(board, action, unused_reward, unused_ovf);
        board = spawn_tile(playing ? moved : board, bits_pos, bits_val);
        fours += ((0x800000u - (bits_val >> 9)) > 7549747u) ? 1u : 0u;
        lm = legal_mask(board);

        if (!playing) {
            phase += 1;  // INIT0 -> INIT1 -> PLAY
            continue;
        }
        ++t;
        const bool done = lm == 0u;
        const bool cut = !done && t >= max_steps;
        if ((t & 255u) == 0u || done || cut) seen15 |= has_max_nibble(board);
        if (done || cut) {
            const uint32_t score = board_potential(board) - 4u * fours;
            if (final_boards) final_boards[e] = board;
            if (lengths) lengths[e] = t;
            if (scores) scores[e] = score;
            if (results) results[e] = make_uint4((uint32_t)board, (uint32_t)(board >> 32), t, score);  // G2048EpisodeResult
            const uint32_t me = max_exponent(board);
            const unsigned long long tile = 1ull << me;
            st_episodes += 1;
            st_steps += t;
            st_score += score;
            st_cut += cut ? 1u : 0u;
            st_ovf += seen15 ? 1u : 0u;
            st_longest = max(st_longest, t);
            st_tile += tile;
            st_tile2 += tile * tile;
            atomicAdd(&s_stats[16 + me], 1ull);
            phase = PHASE_NONE;
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

template <int MODE, int POLICY>
static int launch_play2(const uint32_t* d_subs, int64_t n_subs, int64_t batch_global, int64_t env_lo, int64_t n,
                        uint64_t* d_work, uint64_t* d_final_boards, uint32_t* d_lengths, uint32_t* d_scores,
                        uint64_t* d_stats, cudaStream_t st, void* d_results) {
    const int sms = sm_count();
    if (sms <= 0) return fail_arg("play: no device");
    // A small batch is all latency: one warp per CTA spreads its warps over the SMs and their schedulers -- 1 024 envs
    // are 32 SMs with one warp each instead of 4 SMs with eight, two to a scheduler and close to its issue rate
    // (tools/probes/small_batch_probe.py: 1 024 envs 0.49 -> 0.38 ms, 4 096 envs 0.52 -> 0.38 ms, DRUL 0.70 -> 0.60 ms).
    const int threads = (G2048_PLAY2_NARROW && n <= (int64_t)sms * 4 * 32) ? 32 : PLAY2_THREADS;
    int per_sm = 0;
    int rc = check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, play2_kernel<MODE, POLICY>, threads, 0),
                        "play: occupancy");
    if (rc) return rc;
    if (per_sm <= 0) return fail_arg("play: no device");
    int64_t grid = (int64_t)sms * per_sm;  // one wave of resident CTAs: the kernel is persistent
    const int64_t needed = (n + threads - 1) / threads;
    if (grid > needed) grid = needed;
    play2_kernel<MODE, POLICY><<<(unsigned)grid, threads, 0, st>>>(
        (const uint2*)d_subs, n_subs, (uint32_t)batch_global, (uint32_t)env_lo, (uint32_t)n, (unsigned long long*)d_work,
        (u64*)d_final_boards, d_lengths, d_scores, (unsigned long long*)d_stats, (uint4*)d_results);
    return check_cuda(cudaGetLastError(), "play");
}

}  // namespace g2048

using namespace g2048;

// Batches of at least this many envs go to the table kernel (g2048_play3.cu), whose per-CTA cost of
// copying 192 KiB of tables into shared memory needs a long-running launch to amortise.
constexpr int64_t PLAY_TABLES_MIN_ENVS = 32768;

// the SWAR kernel with every output form (d_results: one 16-byte G2048EpisodeResult per env, may be NULL)
static int play_swar_impl(int policy, const uint32_t* d_subs, int64_t n_subs, int64_t batch_global, int64_t env_lo, int64_t n,
                          int rng_mode, uint64_t* d_work, uint64_t* d_final_boards, uint32_t* d_lengths, uint32_t* d_scores,
                          uint64_t* d_stats, void* d_results, void* stream) {
    G2048_REQUIRE(policy == G2048_POLICY_RANDOM || policy == G2048_POLICY_DRUL, "play: policy");
    G2048_REQUIRE(rng_mode == G2048_RNG_ORIGINAL || rng_mode == G2048_RNG_PARTITIONABLE, "play: rng_mode");
    G2048_REQUIRE(batch_global > 0 && batch_global <= 0x7FFFFFFFll && env_lo >= 0 && n >= 0 && env_lo + n <= batch_global,
                  "play: batch");
    G2048_REQUIRE(n_subs >= 3 && d_subs && d_work && d_stats, "play: pointers");
    G2048_REQUIRE(((uintptr_t)d_results & 15u) == 0, "play: results must be 16-byte aligned");
    if (n == 0) return G2048_OK;
    cudaStream_t st = (cudaStream_t)stream;
#define ARGS d_subs, n_subs, batch_global, env_lo, n, d_work, d_final_boards, d_lengths, d_scores, d_stats, st, d_results
    if (policy == G2048_POLICY_RANDOM) {
        if (rng_mode == G2048_RNG_PARTITIONABLE) return launch_play2<G2048_RNG_PARTITIONABLE, G2048_POLICY_RANDOM>(ARGS);
        return launch_play2<G2048_RNG_ORIGINAL, G2048_POLICY_RANDOM>(ARGS);
    }
    if (rng_mode == G2048_RNG_PARTITIONABLE) return launch_play2<G2048_RNG_PARTITIONABLE, G2048_POLICY_DRUL>(ARGS);
    return launch_play2<G2048_RNG_ORIGINAL, G2048_POLICY_DRUL>(ARGS);
#undef ARGS
}

namespace g2048 {
int play_tables_impl(int policy, const uint32_t* d_subs, int64_t n_subs, int64_t batch_global, int64_t env_lo, int64_t n,
                     int rng_mode, uint64_t* d_work, uint64_t* d_final_boards, uint32_t* d_lengths, uint32_t* d_scores,
                     uint64_t* d_stats, void* d_results, void* stream);  // g2048_play3.cu
}

extern "C" int g2048_play(int policy, const uint32_t* d_subs, int64_t n_subs, int64_t batch_global, int64_t env_lo,
                          int64_t n, int rng_mode, uint64_t* d_work, uint64_t* d_final_boards, uint32_t* d_lengths,
                          uint32_t* d_scores, uint64_t* d_stats, void* stream) {
    if (n >= PLAY_TABLES_MIN_ENVS)
        return g2048_play_tables(policy, d_subs, n_subs, batch_global, env_lo, n, rng_mode, d_work, d_final_boards,
                                 d_lengths, d_scores, d_stats, stream);
    return g2048_play_swar(policy, d_subs, n_subs, batch_global, env_lo, n, rng_mode, d_work, d_final_boards,
                           d_lengths, d_scores, d_stats, stream);
}

// Same run, per-env results as ONE 16-byte record per env (G2048EpisodeResult) instead of three arrays: one store
// per finished episode -- the form g2048_play_host uses when the kernel writes straight into pinned host memory.
extern "C" int g2048_play_packed(int policy, const uint32_t* d_subs, int64_t n_subs, int64_t batch_global, int64_t env_lo,
                                 int64_t n, int rng_mode, uint64_t* d_work, G2048EpisodeResult* d_results, uint64_t* d_stats,
                                 void* stream) {
    if (n >= PLAY_TABLES_MIN_ENVS)
        return g2048::play_tables_impl(policy, d_subs, n_subs, batch_global, env_lo, n, rng_mode, d_work, nullptr, nullptr,
                                       nullptr, d_stats, d_results, stream);
    return play_swar_impl(policy, d_subs, n_subs, batch_global, env_lo, n, rng_mode, d_work, nullptr, nullptr, nullptr,
                          d_stats, d_results, stream);
}

// SWAR kernel entry (no tables, any batch size): same arguments and results as g2048_play.
extern "C" int g2048_play_swar(int policy, const uint32_t* d_subs, int64_t n_subs, int64_t batch_global,
                               int64_t env_lo, int64_t n, int rng_mode, uint64_t* d_work, uint64_t* d_final_boards,
                               uint32_t* d_lengths, uint32_t* d_scores, uint64_t* d_stats, void* stream) {
    return play_swar_impl(policy, d_subs, n_subs, batch_global, env_lo, n, rng_mode, d_work, d_final_boards, d_lengths,
                          d_scores, d_stats, nullptr, stream);
}
