// Per-env device functions: Pgx 2048 init/step semantics and the reference's three action rules,
// written for one thread per env with everything in registers.
#pragma once
#include <cfloat>

#include "../../include/g2048.h"
#include "g2048_board.cuh"

namespace g2048 {

struct EnvState {
    u64 board;
    uint32_t status;  // bits 0-3 legal mask (all set when terminal), 4 done, 5 overflow
};

// Per-env key of a vmapped call: split(*sub, batch_global)[env_lo + i], or -- when batch_global is 0 --
// the caller's explicit key array (n,2), as in jax.vmap(act_fn)(keys, obs, mask).
template <int MODE>
__device__ __forceinline__ Key env_key(const uint32_t* __restrict__ sub, uint32_t batch_global, uint32_t env_lo,
                                       int64_t i) {
    if (batch_global == 0u) return Key{sub[2 * i], sub[2 * i + 1]};
    return split_at<MODE>(Key{sub[0], sub[1]}, batch_global, env_lo + (uint32_t)i);
}

// Pgx _add_random_num(board, key): k1, k2 = split(key); position from k1, value from k2.
template <int MODE>
__device__ __forceinline__ u64 add_random(u64 board, Key k) {
    Key k1, k2;
    split2<MODE>(k, k1, k2);
    return spawn_tile(board, bits_scalar<MODE>(k1), bits_scalar<MODE>(k2));
}

// Pgx 2048 _init(key): r1, r2 = split(key); two spawns on the empty board; exact legal mask.
template <int MODE>
__device__ __forceinline__ EnvState env_init(Key k) {
    Key r1, r2;
    split2<MODE>(k, r1, r2);
    u64 b = add_random<MODE>(0ull, r1);
    b = add_random<MODE>(b, r2);
    return EnvState{b, legal_mask(b)};
}

// pgx core.Env.step + 2048 _step with the spawn draws given.  Returns State.rewards[0].
//   - a finished env is returned unchanged with reward 0;
//   - otherwise move, spawn, recompute the mask; terminated = no legal action;
//   - an action that was illegal under the PRE-step mask: reward -1, terminated (the no-op move
//     and the spawn have still been applied);
//   - a terminal state's mask is all-True.
__device__ __forceinline__ float env_step_draws(EnvState& s, int action, uint32_t bits_pos, uint32_t bits_val) {
    if (s.status & G2048_STATUS_DONE) return 0.0f;
    const bool illegal = ((s.status >> action) & 1u) == 0u;
    uint32_t reward = 0;
    bool overflow = false;
    u64 b = move_board(s.board, action, reward, overflow);
    b = spawn_tile(b, bits_pos, bits_val);
    uint32_t lm = legal_mask(b);
    const bool term = (lm == 0u) | illegal;
    if (term) lm = G2048_STATUS_MASK;
    s.board = b;
    s.status = lm | (term ? G2048_STATUS_DONE : 0u) | ((overflow || (s.status & G2048_STATUS_OVERFLOW)) ? G2048_STATUS_OVERFLOW : 0u);
    return illegal ? -1.0f : (float)reward;
}

template <int MODE>
__device__ __forceinline__ float env_step(EnvState& s, int action, Key step_key) {
    if (s.status & G2048_STATUS_DONE) return 0.0f;
    Key k1, k2;
    split2<MODE>(step_key, k1, k2);
    return env_step_draws(s, action, bits_scalar<MODE>(k1), bits_scalar<MODE>(k2));
}

// act_randomly (src/actions/act_randomly.py:40-48): categorical over logits that are equal on the
// legal actions and -FLT_MAX elsewhere.  argmax(gumbel(u_i) + c) over the legal i is argmax u_i,
// and u_i is monotone in (bits_i >> 9), so the draw needs no float at all.  First index wins ties
// (argmax).  No legal action (never the case on a live env) -> uniform over all four.
template <int MODE>
__device__ __forceinline__ int act_random(Key k, uint32_t legal) {
    uint32_t bits[4];
    bits4<MODE>(k, bits);
    const uint32_t allowed = (legal & 15u) ? (legal & 15u) : 15u;
    int best = 0;
    int best_v = -1;
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

// log(probs[action]) with probs = mask / n (act_randomly.py:40-44,51)
__device__ __forceinline__ float act_random_log_prob(uint32_t legal) {
    const int n = __popc(legal & 15u);
    return logf(n > 0 ? __fdiv_rn(1.0f, (float)n) : 0.25f);
}

// act_drul (src/actions/act_drul.py:40-44): first legal of Down, Right, Up, Left; none -> Down.
__device__ __forceinline__ int act_drul(uint32_t legal) {
    const uint32_t m = legal & 15u;
    return m ? (31 - __clz((int)m)) : 3;
}

// TorchActionFunction.__call__ after the network (src/ppo/torch_action_wrapper.py:84-102):
// clip, categorical (gumbel-max on jax's draws) or argmax, log_prob = logit[a] - logsumexp.
// The mask rule is PPOAgent.forward's (src/ppo/ppo_agent.py:117-121): logits - 1e8 * (1 - mask).
struct Logits4 {
    float v[4];
};

__device__ __forceinline__ Logits4 prepare_logits(const float4 raw, uint32_t legal, bool use_mask) {
    Logits4 l{{raw.x, raw.y, raw.z, raw.w}};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (use_mask) l.v[i] = __fsub_rn(l.v[i], __fmul_rn(1e8f, ((legal >> i) & 1u) ? 0.0f : 1.0f));
        l.v[i] = fmaxf(l.v[i], -FLT_MAX);
    }
    return l;
}

__device__ __forceinline__ float log_sum_exp4(const Logits4& l) {
    const float m = fmaxf(fmaxf(l.v[0], l.v[1]), fmaxf(l.v[2], l.v[3]));
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += expf(l.v[i] - m);
    return m + logf(s);
}

template <int MODE>
__device__ __forceinline__ int sample_categorical(Key k, const Logits4& l) {
    uint32_t bits[4];
    bits4<MODE>(k, bits);
    int best = 0;
    float best_v = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float u = unit_float_tiny(bits[i]);
        const float g = -logf(-logf(u));
        const float v = g + l.v[i];
        if (i == 0 || v > best_v) {
            best_v = v;
            best = i;
        }
    }
    return best;
}

__device__ __forceinline__ int argmax4(const Logits4& l) {
    int best = 0;
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (l.v[i] > l.v[best]) best = i;
    return best;
}

// Categorical(logits).entropy() as torch computes it: -sum p * logp with logp clamped at FLT_MIN log
__device__ __forceinline__ float entropy4(const Logits4& l, float lse) {
    float h = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float lp = fmaxf(l.v[i] - lse, -FLT_MAX);
        h -= expf(lp) * lp;
    }
    return h;
}

}  // namespace g2048
