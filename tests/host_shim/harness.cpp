// TEST-ONLY: compiles the product's per-env device functions for the host (see cuda_shim.h) and
// exposes batched loops over them for tests/test_device_logic_host.py.
#include "cuda_shim.h"
#include "../../2048-ppo-agent_b200/csrc/g2048_env.cuh"

using namespace g2048;

template <int MODE>
static void init_t(const uint32_t* sub, uint32_t batch, uint32_t lo, int64_t n, uint64_t* boards, uint8_t* status) {
    for (int64_t i = 0; i < n; ++i) {
        EnvState s = env_init<MODE>(split_at<MODE>(Key{sub[0], sub[1]}, batch, lo + (uint32_t)i));
        boards[i] = s.board;
        status[i] = (uint8_t)s.status;
    }
}

template <int MODE>
static void step_t(uint64_t* boards, uint8_t* status, const int32_t* actions, const uint32_t* sub, uint32_t batch,
                   uint32_t lo, int64_t n, float* rewards) {
    for (int64_t i = 0; i < n; ++i) {
        EnvState s{boards[i], status[i]};
        rewards[i] = env_step<MODE>(s, actions[i] & 3, split_at<MODE>(Key{sub[0], sub[1]}, batch, lo + (uint32_t)i));
        boards[i] = s.board;
        status[i] = (uint8_t)s.status;
    }
}

template <int MODE>
static void act_t(int policy, const uint8_t* status, const uint32_t* sub, uint32_t batch, uint32_t lo, int64_t n,
                  int32_t* actions, float* log_probs) {
    for (int64_t i = 0; i < n; ++i) {
        const uint32_t lm = status[i] & 15u;
        if (policy == 0) {
            actions[i] = act_random<MODE>(split_at<MODE>(Key{sub[0], sub[1]}, batch, lo + (uint32_t)i), lm);
            log_probs[i] = act_random_log_prob(lm);
        } else {
            actions[i] = act_drul(lm);
        }
    }
}

template <int MODE>
static void sample_t(const float* logits, const uint8_t* status, int use_mask, int sample, const uint32_t* sub,
                     uint32_t batch, uint32_t lo, int64_t n, int32_t* actions, float* log_probs, float* entropy) {
    for (int64_t i = 0; i < n; ++i) {
        const float4 raw{logits[4 * i], logits[4 * i + 1], logits[4 * i + 2], logits[4 * i + 3]};
        const Logits4 l = prepare_logits(raw, status[i] & 15u, use_mask != 0);
        const int a = sample ? sample_categorical<MODE>(split_at<MODE>(Key{sub[0], sub[1]}, batch, lo + (uint32_t)i), l)
                             : argmax4(l);
        const float lse = log_sum_exp4(l);
        actions[i] = a;
        log_probs[i] = l.v[a] - lse;
        entropy[i] = entropy4(l, lse);
    }
}

extern "C" {
// every 16-cell flag mask x every valid k: the cell kth_flag_cell picks
void shim_kth_flag_cell_all(int32_t* out /* [65536][16], -1 where k exceeds the number of flags */) {
    for (uint32_t mask = 0; mask < 65536u; ++mask) {
        unsigned long long flags = 0;
        for (int i = 0; i < 16; ++i) if (mask >> i & 1u) flags |= 1ull << (4 * i);
        const int count = __builtin_popcount(mask);
        for (int k = 1; k <= 16; ++k) out[mask * 16 + (k - 1)] = k <= count ? kth_flag_cell(flags, k) : -1;
    }
}
void shim_threefry(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t* out) {
    Key y = threefry2x32(Key{k0, k1}, x0, x1);
    out[0] = y.a;
    out[1] = y.b;
}
void shim_split(const uint32_t* key, uint32_t n, int mode, uint32_t* out) {
    for (uint32_t i = 0; i < n; ++i) {
        Key k = mode ? split_at<1>(Key{key[0], key[1]}, n, i) : split_at<0>(Key{key[0], key[1]}, n, i);
        out[2 * i] = k.a;
        out[2 * i + 1] = k.b;
    }
}
void shim_env_init(const uint32_t* sub, uint32_t batch, uint32_t lo, int64_t n, int mode, uint64_t* boards, uint8_t* status) {
    if (mode) init_t<1>(sub, batch, lo, n, boards, status); else init_t<0>(sub, batch, lo, n, boards, status);
}
void shim_env_step(uint64_t* boards, uint8_t* status, const int32_t* actions, const uint32_t* sub, uint32_t batch,
                   uint32_t lo, int64_t n, int mode, float* rewards) {
    if (mode) step_t<1>(boards, status, actions, sub, batch, lo, n, rewards);
    else step_t<0>(boards, status, actions, sub, batch, lo, n, rewards);
}
void shim_env_step_draws(uint64_t* boards, uint8_t* status, const int32_t* actions, const uint32_t* bp,
                         const uint32_t* bv, int64_t n, float* rewards) {
    for (int64_t i = 0; i < n; ++i) {
        EnvState s{boards[i], status[i]};
        rewards[i] = env_step_draws(s, actions[i] & 3, bp[i], bv[i]);
        boards[i] = s.board;
        status[i] = (uint8_t)s.status;
    }
}
void shim_act(int policy, const uint8_t* status, const uint32_t* sub, uint32_t batch, uint32_t lo, int64_t n, int mode,
              int32_t* actions, float* log_probs) {
    if (mode) act_t<1>(policy, status, sub, batch, lo, n, actions, log_probs);
    else act_t<0>(policy, status, sub, batch, lo, n, actions, log_probs);
}
void shim_sample(const float* logits, const uint8_t* status, int use_mask, int sample, const uint32_t* sub,
                 uint32_t batch, uint32_t lo, int64_t n, int mode, int32_t* actions, float* log_probs, float* entropy) {
    if (mode) sample_t<1>(logits, status, use_mask, sample, sub, batch, lo, n, actions, log_probs, entropy);
    else sample_t<0>(logits, status, use_mask, sample, sub, batch, lo, n, actions, log_probs, entropy);
}
void shim_move(const uint64_t* boards, const int32_t* actions, int64_t n, uint64_t* out, uint32_t* rewards, uint8_t* masks) {
    for (int64_t i = 0; i < n; ++i) {
        bool ovf = false;
        out[i] = move_board(boards[i], actions[i] & 3, rewards[i], ovf);
        masks[i] = (uint8_t)(legal_mask(boards[i]) | (ovf ? 0x20 : 0));
    }
}
// the keyed bijection behind g2048_random_subset: positions P(first), ..., P(first + m - 1) of [0, n)
void shim_random_subset(uint32_t k0, uint32_t k1, uint64_t n, uint64_t first, int64_t m, int64_t* out) {
    const int h = feistel_half_bits(n);
    for (int64_t i = 0; i < m; ++i) out[i] = (int64_t)feistel_position(Key{k0, k1}, first + (uint64_t)i, h, n);
}
}
