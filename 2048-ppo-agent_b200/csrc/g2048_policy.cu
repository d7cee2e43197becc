// Policy-logit kernels: masked categorical sampling / argmax with jax-compatible draws, log-prob,
// entropy, and the fused "sample -> env.step -> rollout record" loop step that replaces one
// iteration of src/runs/batch_runner.py:117-136 when the policy is a network
// (src/ppo/torch_action_wrapper.py:71-104).  The network forward itself stays PyTorch/cuBLAS.
#include "g2048_common.cuh"
#include "g2048_env.cuh"

namespace g2048 {

__device__ __forceinline__ float pick4(const Logits4& l, int a) {
    return a == 0 ? l.v[0] : (a == 1 ? l.v[1] : (a == 2 ? l.v[2] : l.v[3]));
}

template <int MODE>
__global__ void __launch_bounds__(256)
policy_step_kernel(u64* __restrict__ boards, uint8_t* __restrict__ status, const float4* __restrict__ logits,
                   const float* __restrict__ values, int use_mask, int sample, int auto_reset,
                   const uint32_t* __restrict__ sub_act, const uint32_t* __restrict__ sub_step, uint32_t batch_global,
                   uint32_t env_lo, int64_t n, u64* __restrict__ rec_boards, uint8_t* __restrict__ rec_meta,
                   float* __restrict__ rec_rewards, float* __restrict__ rec_log_probs, float* __restrict__ rec_values,
                   int32_t* __restrict__ actions_out, const int32_t* __restrict__ step_index,
                   const int64_t* __restrict__ env_ids, int64_t n_envs) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    // live-env compaction: thread `row` serves env env_ids[row] of the shard; logits / values are compact (one row per
    // live env), state, records and the RNG counter use the env's own index
    const int64_t i = env_ids ? env_ids[row] : row;
    if (step_index) {
        // graph-replay form: the step number lives in device memory; sub keys of step t sit 4 words further per
        // step (act key, step key), the record slot n entries further
        const int64_t t = *step_index;
        sub_act += 4 * t;
        sub_step += 4 * t;
        const int64_t at = t * n_envs;  // record rows are n_envs long whatever the number of live rows
        if (rec_boards) rec_boards += at;
        if (rec_meta) rec_meta += at;
        if (rec_rewards) rec_rewards += at;
        if (rec_log_probs) rec_log_probs += at;
        if (rec_values) rec_values += at;
    }
    EnvState s{boards[i], status[i]};
    // live-env list that was built a few steps ago: an env that has finished since then is skipped (its state is frozen
    // and the record slots after its end stay untouched, as if it had left the list)
    if (env_ids && !auto_reset && (s.status & G2048_STATUS_DONE)) return;
    // pgx.experimental.auto_reset: a state that finished on the previous step was already replaced
    // by a fresh one but still carries terminated=True; the wrapper clears it before stepping.
    if (auto_reset && (s.status & G2048_STATUS_DONE)) s.status &= ~(uint32_t)G2048_STATUS_DONE;
    const uint32_t lm = s.status & G2048_STATUS_MASK;
    const Logits4 l = prepare_logits(logits[row], lm, use_mask != 0);
    int a;
    if (sample) {
        a = sample_categorical<MODE>(env_key<MODE>(sub_act, batch_global, env_lo, i), l);
    } else {
        a = argmax4(l);
    }
    const float lp = pick4(l, a) - log_sum_exp4(l);
    const u64 pre = s.board;
    const Key step_key = env_key<MODE>(sub_step, batch_global, env_lo, i);
    float r;
    if (auto_reset) {
        Key k1, k2;
        split2<MODE>(step_key, k1, k2);
        r = env_step<MODE>(s, a, k1);
        if (s.status & G2048_STATUS_DONE) {
            const EnvState fresh = env_init<MODE>(k2);
            s.board = fresh.board;
            s.status = fresh.status | G2048_STATUS_DONE | (s.status & G2048_STATUS_OVERFLOW);
        }
    } else {
        r = env_step<MODE>(s, a, step_key);
    }
    boards[i] = s.board;
    status[i] = (uint8_t)s.status;
    const bool done = (s.status & G2048_STATUS_DONE) != 0u;
    if (rec_boards) rec_boards[i] = pre;
    if (rec_meta) rec_meta[i] = (uint8_t)((uint32_t)a | (lm << 2) | (done ? 0x40u : 0u));
    if (rec_rewards) rec_rewards[i] = r;
    if (rec_log_probs) rec_log_probs[i] = lp;
    if (rec_values && values) rec_values[i] = values[row];
    if (actions_out) actions_out[i] = a;
}

__global__ void counter_add_kernel(int32_t* counter, int32_t delta) { *counter += delta; }

template <int MODE>
__global__ void __launch_bounds__(256)
sample_logits_kernel(const float4* __restrict__ logits, const uint8_t* __restrict__ status, int use_mask, int sample,
                     const uint32_t* __restrict__ sub_act, uint32_t batch_global, uint32_t env_lo, int64_t n,
                     int32_t* __restrict__ actions, float* __restrict__ log_probs, float* __restrict__ entropy) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t lm = status ? (status[i] & G2048_STATUS_MASK) : 15u;
    const Logits4 l = prepare_logits(logits[i], lm, use_mask != 0);
    int a;
    if (sample) {
        a = sample_categorical<MODE>(env_key<MODE>(sub_act, batch_global, env_lo, i), l);
    } else {
        a = argmax4(l);
    }
    const float lse = log_sum_exp4(l);
    actions[i] = a;
    if (log_probs) log_probs[i] = pick4(l, a) - lse;
    if (entropy) entropy[i] = entropy4(l, lse);
}

__global__ void __launch_bounds__(256)
evaluate_logits_kernel(const float4* __restrict__ logits, const uint8_t* __restrict__ mask_bits, int use_mask,
                       const int32_t* __restrict__ actions, int64_t n, float* __restrict__ log_probs,
                       float* __restrict__ entropy) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t lm = mask_bits ? (mask_bits[i] & G2048_STATUS_MASK) : 15u;
    const Logits4 l = prepare_logits(logits[i], lm, use_mask != 0);
    const float lse = log_sum_exp4(l);
    if (log_probs) log_probs[i] = pick4(l, actions[i] & 3) - lse;
    if (entropy) entropy[i] = entropy4(l, lse);
}

}  // namespace g2048

using namespace g2048;

static inline bool valid_mode(int m) { return m == G2048_RNG_ORIGINAL || m == G2048_RNG_PARTITIONABLE; }
static inline bool valid_batch(int64_t batch_global, int64_t env_lo, int64_t n) {
    if (batch_global == 0) return env_lo == 0 && n >= 0;  // explicit (n,2) key arrays
    return batch_global > 0 && batch_global <= 0x7FFFFFFFll && env_lo >= 0 && n >= 0 && env_lo + n <= batch_global;
}
static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

static int launch_policy_step(uint64_t* d_boards, uint8_t* d_status, const float* d_logits, const float* d_values,
                              int use_mask, int sample, int auto_reset, const uint32_t* d_sub_act,
                              const uint32_t* d_sub_step, const int32_t* d_step_index, const int64_t* d_env_ids,
                              int64_t n_rows, int64_t batch_global, int64_t env_lo, int64_t n, int rng_mode, uint64_t* d_rec_boards, uint8_t* d_rec_meta,
                              float* d_rec_rewards, float* d_rec_log_probs, float* d_rec_values, int32_t* d_actions_out,
                              void* stream) {
    G2048_REQUIRE(valid_mode(rng_mode) && valid_batch(batch_global, env_lo, n), "policy_step: batch");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_boards && d_status && d_logits && d_sub_step && (d_sub_act || !sample), "policy_step: pointers");
    G2048_REQUIRE(aligned16(d_logits), "policy_step: logits must be 16-byte aligned (n,4) float32");
    const uint32_t* sub_act = d_sub_act ? d_sub_act : d_sub_step;
    if (d_env_ids) {
        G2048_REQUIRE(n_rows >= 0 && n_rows <= n, "policy_step: more rows than envs");
        if (n_rows == 0) return G2048_OK;
    } else {
        n_rows = n;
    }
    const unsigned g = blocks_for(n_rows, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (rng_mode == G2048_RNG_PARTITIONABLE) {
        policy_step_kernel<G2048_RNG_PARTITIONABLE><<<g, 256, 0, st>>>(
            (u64*)d_boards, d_status, (const float4*)d_logits, d_values, use_mask, sample, auto_reset, sub_act,
            d_sub_step, (uint32_t)batch_global, (uint32_t)env_lo, n_rows, (u64*)d_rec_boards, d_rec_meta, d_rec_rewards,
            d_rec_log_probs, d_rec_values, d_actions_out, d_step_index, d_env_ids, n);
    } else {
        policy_step_kernel<G2048_RNG_ORIGINAL><<<g, 256, 0, st>>>(
            (u64*)d_boards, d_status, (const float4*)d_logits, d_values, use_mask, sample, auto_reset, sub_act,
            d_sub_step, (uint32_t)batch_global, (uint32_t)env_lo, n_rows, (u64*)d_rec_boards, d_rec_meta, d_rec_rewards,
            d_rec_log_probs, d_rec_values, d_actions_out, d_step_index, d_env_ids, n);
    }
    G2048_CHECK_LAUNCH("policy_step");
    return G2048_OK;
}

extern "C" int g2048_policy_step(uint64_t* d_boards, uint8_t* d_status, const float* d_logits, const float* d_values,
                                 int use_mask, int sample, int auto_reset, const uint32_t* d_sub_act,
                                 const uint32_t* d_sub_step, int64_t batch_global, int64_t env_lo, int64_t n,
                                 int rng_mode, uint64_t* d_rec_boards, uint8_t* d_rec_meta, float* d_rec_rewards,
                                 float* d_rec_log_probs, float* d_rec_values, int32_t* d_actions_out, void* stream) {
    return launch_policy_step(d_boards, d_status, d_logits, d_values, use_mask, sample, auto_reset, d_sub_act, d_sub_step,
                              nullptr, nullptr, 0, batch_global, env_lo, n, rng_mode, d_rec_boards, d_rec_meta, d_rec_rewards,
                              d_rec_log_probs, d_rec_values, d_actions_out, stream);
}

extern "C" int g2048_policy_step_live(uint64_t* d_boards, uint8_t* d_status, const float* d_logits, const float* d_values,
                                      int use_mask, int sample, int auto_reset, const uint32_t* d_sub_act,
                                      const uint32_t* d_sub_step, const int64_t* d_env_ids, int64_t n_live,
                                      int64_t batch_global, int64_t env_lo, int64_t n, int rng_mode, uint64_t* d_rec_boards,
                                      uint8_t* d_rec_meta, float* d_rec_rewards, float* d_rec_log_probs, float* d_rec_values,
                                      int32_t* d_actions_out, void* stream) {
    G2048_REQUIRE(d_env_ids || n_live == 0, "policy_step_live: env ids");
    if (n_live == 0) return G2048_OK;
    return launch_policy_step(d_boards, d_status, d_logits, d_values, use_mask, sample, auto_reset, d_sub_act, d_sub_step,
                              nullptr, d_env_ids, n_live, batch_global, env_lo, n, rng_mode, d_rec_boards, d_rec_meta,
                              d_rec_rewards, d_rec_log_probs, d_rec_values, d_actions_out, stream);
}

extern "C" int g2048_policy_step_at(uint64_t* d_boards, uint8_t* d_status, const float* d_logits, const float* d_values,
                                    int use_mask, int sample, int auto_reset, const uint32_t* d_subs,
                                    const int32_t* d_step_index, int64_t batch_global, int64_t env_lo, int64_t n,
                                    int rng_mode, uint64_t* d_rec_boards, uint8_t* d_rec_meta, float* d_rec_rewards,
                                    float* d_rec_log_probs, float* d_rec_values, int32_t* d_actions_out, void* stream) {
    G2048_REQUIRE(d_subs && d_step_index, "policy_step_at: sub keys and step index");
    return launch_policy_step(d_boards, d_status, d_logits, d_values, use_mask, sample, auto_reset, d_subs, d_subs + 2,
                              d_step_index, nullptr, 0, batch_global, env_lo, n, rng_mode, d_rec_boards, d_rec_meta, d_rec_rewards,
                              d_rec_log_probs, d_rec_values, d_actions_out, stream);
}

extern "C" int g2048_counter_add(int32_t* d_counter, int32_t delta, void* stream) {
    G2048_REQUIRE(d_counter, "counter_add: pointer");
    counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(d_counter, delta);
    G2048_CHECK_LAUNCH("counter_add");
    return G2048_OK;
}

extern "C" int g2048_sample_logits(const float* d_logits, const uint8_t* d_status, int use_mask, int sample,
                                   const uint32_t* d_sub_act, int64_t batch_global, int64_t env_lo, int64_t n,
                                   int rng_mode, int32_t* d_actions, float* d_log_probs, float* d_entropy,
                                   void* stream) {
    G2048_REQUIRE(valid_mode(rng_mode) && valid_batch(batch_global, env_lo, n), "sample_logits: batch");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_logits && d_actions && (d_sub_act || !sample) && (d_status || !use_mask), "sample_logits: pointers");
    G2048_REQUIRE(aligned16(d_logits), "sample_logits: logits must be 16-byte aligned (n,4) float32");
    const uint32_t dummy_ok = 0;
    (void)dummy_ok;
    const unsigned g = blocks_for(n, 256);
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t* sub = d_sub_act ? d_sub_act : (const uint32_t*)d_logits;  // never dereferenced when !sample
    if (rng_mode == G2048_RNG_PARTITIONABLE) {
        sample_logits_kernel<G2048_RNG_PARTITIONABLE><<<g, 256, 0, st>>>((const float4*)d_logits, d_status, use_mask,
            sample, sub, (uint32_t)batch_global, (uint32_t)env_lo, n, d_actions, d_log_probs, d_entropy);
    } else {
        sample_logits_kernel<G2048_RNG_ORIGINAL><<<g, 256, 0, st>>>((const float4*)d_logits, d_status, use_mask,
            sample, sub, (uint32_t)batch_global, (uint32_t)env_lo, n, d_actions, d_log_probs, d_entropy);
    }
    G2048_CHECK_LAUNCH("sample_logits");
    return G2048_OK;
}

extern "C" int g2048_evaluate_logits(const float* d_logits, const uint8_t* d_mask_bits, int use_mask,
                                     const int32_t* d_actions, int64_t n, float* d_log_probs, float* d_entropy,
                                     void* stream) {
    G2048_REQUIRE(n >= 0, "evaluate_logits: n");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_logits && d_actions && (d_mask_bits || !use_mask), "evaluate_logits: pointers");
    G2048_REQUIRE(aligned16(d_logits), "evaluate_logits: logits must be 16-byte aligned (n,4) float32");
    evaluate_logits_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)d_logits, d_mask_bits, use_mask, d_actions, n, d_log_probs, d_entropy);
    G2048_CHECK_LAUNCH("evaluate_logits");
    return G2048_OK;
}
