/* CPU restatement (plain C) of the reference's rollout hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * build, load or call this file; the product (libg2048.so and the g2048 package) never does.
 *
 * Parity status: PINNED -- this file is checked (tests/test_oracle.py) against the golden
 * trajectories and histograms decoded from the reference's own Pgx/JAX artefacts
 * (tests/golden/svg_trajectories.npz, tests/golden/histograms.json) and against the numpy
 * restatement in pgx2048_oracle.py.  It keeps boards as 16 small integers and walks rows
 * cell by cell -- deliberately nothing like the bitboard CUDA path it checks.
 *
 * Follows (paths relative to the reference repo):
 *   jax.random on Threefry-2x32 (jax==0.5.3, uv.lock:701-702), both counter layouts;
 *   Pgx "2048" init/step (pgx==2.6.0, uv.lock:1564-1565) as called from
 *     src/runs/batch_runner.py:105-136 and src/runs/run_actions_batch.py:41-55;
 *   src/actions/act_randomly.py:40-51, src/actions/act_drul.py:40-44;
 *   src/ppo/data_loader.py:103-130 (GAE).
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared -fPIC; -ffp-contract=off so that no
 * fused multiply-add changes the fp32 results).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MODE_ORIGINAL 0
#define MODE_PARTITIONABLE 1
#define POLICY_RANDOM 0
#define POLICY_DRUL 1

typedef struct { uint32_t a, b; } key_t2;

static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

static key_t2 threefry(key_t2 k, uint32_t x0, uint32_t x1) {
    static const int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
    uint32_t ks[3] = {k.a, k.b, k.a ^ k.b ^ 0x1BD11BDAu};
    x0 += ks[0];
    x1 += ks[1];
    for (int i = 0; i < 5; ++i) {
        for (int j = 0; j < 4; ++j) {
            x0 += x1;
            x1 = rotl32(x1, R[i % 2][j]);
            x1 ^= x0;
        }
        x0 += ks[(i + 1) % 3];
        x1 += ks[(i + 2) % 3] + (uint32_t)(i + 1);
    }
    key_t2 out = {x0, x1};
    return out;
}

void orc_threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t* out) {
    key_t2 k = {k0, k1};
    key_t2 y = threefry(k, x0, x1);
    out[0] = y.a;
    out[1] = y.b;
}

/* word m of the flattened (n,2) output of the original-layout split(key, n) */
static uint32_t split_word_original(key_t2 k, uint32_t n, uint32_t m) {
    if (m < n) return threefry(k, m, n + m).a;
    return threefry(k, m - n, m).b;
}

/* jax.random.split(key, n)[i] */
static key_t2 split_at(key_t2 k, uint32_t n, uint32_t i, int mode) {
    key_t2 out;
    if (mode == MODE_PARTITIONABLE) return threefry(k, 0u, i);
    out.a = split_word_original(k, n, 2 * i);
    out.b = split_word_original(k, n, 2 * i + 1);
    return out;
}

void orc_split(const uint32_t* key, uint32_t n, int mode, uint32_t* out) {
    key_t2 k = {key[0], key[1]};
    for (uint32_t i = 0; i < n; ++i) {
        key_t2 s = split_at(k, n, i, mode);
        out[2 * i] = s.a;
        out[2 * i + 1] = s.b;
    }
}

/* element i of random_bits(key, shape (m,)); m == 1 is the scalar shape */
static uint32_t bits_at(key_t2 k, uint32_t m, uint32_t i, int mode) {
    if (mode == MODE_PARTITIONABLE) {
        key_t2 y = threefry(k, 0u, i);
        return y.a ^ y.b;
    }
    uint32_t padded = m + (m & 1u);
    uint32_t h = padded / 2;
    if (i < h) {
        uint32_t hi = h + i;
        return threefry(k, i, (hi < m) ? hi : 0u).a;
    }
    return threefry(k, i - h, i).b;
}

static float unit_float(uint32_t bits) {
    union { uint32_t u; float f; } v;
    v.u = (bits >> 9) | 0x3F800000u;
    return v.f - 1.0f;
}

static float uniform01(key_t2 k, int mode) { return unit_float(bits_at(k, 1, 0, mode)); }

/* uniform(key, (4,), minval=tiny, maxval=1) */
static void uniform4_tiny(key_t2 k, int mode, float* u) {
    const float tiny = 1.17549435e-38f;
    for (uint32_t i = 0; i < 4; ++i) {
        float f = unit_float(bits_at(k, 4, i, mode));
        float v = f * (1.0f - tiny) + tiny;
        u[i] = v > tiny ? v : tiny;
    }
}

/* ------------------------------------------------------------------ Pgx 2048 */
static int row_left(uint8_t* row) { /* in place; returns reward */
    uint8_t tiles[4];
    int n = 0, reward = 0, w = 0;
    for (int c = 0; c < 4; ++c)
        if (row[c]) tiles[n++] = row[c];
    memset(row, 0, 4);
    for (int j = 0; j < n;) {
        if (j + 1 < n && tiles[j] == tiles[j + 1]) {
            row[w++] = (uint8_t)(tiles[j] + 1);
            reward += 1 << (tiles[j] + 1);
            j += 2;
        } else {
            row[w++] = tiles[j];
            j += 1;
        }
    }
    return reward;
}

/* 0=Left 1=Up 2=Right 3=Down; returns reward */
static int move_board(uint8_t* b, int action) {
    int reward = 0;
    for (int line = 0; line < 4; ++line) {
        int idx[4];
        for (int k = 0; k < 4; ++k) {
            switch (action) {
                case 0: idx[k] = 4 * line + k; break;
                case 2: idx[k] = 4 * line + (3 - k); break;
                case 1: idx[k] = 4 * k + line; break;
                default: idx[k] = 4 * (3 - k) + line; break;
            }
        }
        uint8_t row[4] = {b[idx[0]], b[idx[1]], b[idx[2]], b[idx[3]]};
        reward += row_left(row);
        for (int k = 0; k < 4; ++k) b[idx[k]] = row[k];
    }
    return reward;
}

static void exact_legal(const uint8_t* b, uint8_t* mask) {
    for (int a = 0; a < 4; ++a) {
        uint8_t t[16];
        memcpy(t, b, 16);
        move_board(t, a);
        mask[a] = memcmp(t, b, 16) != 0;
    }
}

static void add_random_given(uint8_t* b, float u_pos, float u_val) {
    float cum[16], c = 0.0f;
    for (int i = 0; i < 16; ++i) {
        c += (b[i] == 0) ? 1.0f : 0.0f;
        cum[i] = c;
    }
    float r = cum[15] * (1.0f - u_pos);
    int pos = 0;
    while (pos < 15 && cum[pos] < r) ++pos; /* searchsorted side=left */
    float c0 = 0.9f, c1 = c0 + 0.1f;
    float rv = c1 * (1.0f - u_val);
    int val = 1 + (c0 < rv) + (c1 < rv);
    b[pos] = (uint8_t)val;
}

static void add_random(uint8_t* b, key_t2 k, int mode) {
    key_t2 k1 = split_at(k, 2, 0, mode), k2 = split_at(k, 2, 1, mode);
    add_random_given(b, uniform01(k1, mode), uniform01(k2, mode));
}

static void init_one(key_t2 k, int mode, uint8_t* b, uint8_t* mask) {
    key_t2 r1 = split_at(k, 2, 0, mode), r2 = split_at(k, 2, 1, mode);
    memset(b, 0, 16);
    add_random(b, r1, mode);
    add_random(b, r2, mode);
    exact_legal(b, mask);
}

/* returns reward; updates board, mask, done in place */
static float step_one_given(uint8_t* b, uint8_t* mask, uint8_t* done, int action, float u_pos, float u_val) {
    if (*done) return 0.0f; /* frozen */
    int illegal = !mask[action];
    float reward = (float)move_board(b, action);
    add_random_given(b, u_pos, u_val);
    exact_legal(b, mask);
    int term = !(mask[0] | mask[1] | mask[2] | mask[3]);
    if (illegal) {
        reward = -1.0f;
        term = 1;
    }
    if (term) mask[0] = mask[1] = mask[2] = mask[3] = 1;
    *done = (uint8_t)term;
    return reward;
}

static float step_one(uint8_t* b, uint8_t* mask, uint8_t* done, int action, key_t2 k, int mode) {
    if (*done) return 0.0f;
    key_t2 k1 = split_at(k, 2, 0, mode), k2 = split_at(k, 2, 1, mode);
    return step_one_given(b, mask, done, action, uniform01(k1, mode), uniform01(k2, mode));
}

static int act_random_one(key_t2 k, const uint8_t* mask, int mode, float* log_prob) {
    int n = mask[0] + mask[1] + mask[2] + mask[3];
    float probs[4], u[4];
    uniform4_tiny(k, mode, u);
    int best = 0;
    float bestv = 0.0f;
    for (int a = 0; a < 4; ++a) {
        probs[a] = n > 0 ? (float)mask[a] / (float)n : 0.25f;
        float logit = probs[a] > 0.0f ? logf(probs[a]) : -INFINITY;
        if (logit < -3.40282347e+38f) logit = -3.40282347e+38f;
        float g = -logf(-logf(u[a]));
        float v = g + logit;
        if (a == 0 || v > bestv) {
            best = a;
            bestv = v;
        }
    }
    if (log_prob) *log_prob = logf(probs[best]);
    return best;
}

static int act_drul_one(const uint8_t* mask) {
    static const int order[4] = {3, 2, 1, 0};
    for (int i = 0; i < 4; ++i)
        if (mask[order[i]]) return order[i];
    return 3;
}

/* ------------------------------------------------------------------ batched entry points */
void orc_env_init(const uint32_t* keys, int64_t n, int mode, uint8_t* boards, uint8_t* masks) {
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < n; ++e) {
        key_t2 k = {keys[2 * e], keys[2 * e + 1]};
        init_one(k, mode, boards + 16 * e, masks + 4 * e);
    }
}

void orc_spawn_draws(const uint32_t* keys, int64_t n, int mode, float* u_pos, float* u_val) {
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < n; ++e) {
        key_t2 k = {keys[2 * e], keys[2 * e + 1]};
        u_pos[e] = uniform01(split_at(k, 2, 0, mode), mode);
        u_val[e] = uniform01(split_at(k, 2, 1, mode), mode);
    }
}

void orc_env_step_given(uint8_t* boards, uint8_t* masks, uint8_t* done, const int32_t* actions,
                        const float* u_pos, const float* u_val, int64_t n, float* rewards) {
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < n; ++e)
        rewards[e] = step_one_given(boards + 16 * e, masks + 4 * e, done + e, actions[e], u_pos[e], u_val[e]);
}

void orc_env_step(uint8_t* boards, uint8_t* masks, uint8_t* done, const int32_t* actions,
                  const uint32_t* keys, int64_t n, int mode, float* rewards) {
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < n; ++e) {
        key_t2 k = {keys[2 * e], keys[2 * e + 1]};
        rewards[e] = step_one(boards + 16 * e, masks + 4 * e, done + e, actions[e], k, mode);
    }
}

void orc_act(const uint32_t* keys, const uint8_t* masks, int64_t n, int policy, int mode,
             int32_t* actions, float* log_probs) {
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < n; ++e) {
        if (policy == POLICY_RANDOM) {
            key_t2 k = {keys[2 * e], keys[2 * e + 1]};
            actions[e] = act_random_one(k, masks + 4 * e, mode, log_probs ? log_probs + e : NULL);
        } else {
            actions[e] = act_drul_one(masks + 4 * e);
        }
    }
}

/* The runner's chain (batch_runner.py:105,118,126): key, sub = split(key).  Writes n_sub
 * successive sub keys and the advanced chain key back into key_io. */
void orc_chain(uint32_t* key_io, int mode, int64_t n_sub, uint32_t* subs) {
    key_t2 k = {key_io[0], key_io[1]};
    for (int64_t i = 0; i < n_sub; ++i) {
        key_t2 nk = split_at(k, 2, 0, mode), sub = split_at(k, 2, 1, mode);
        subs[2 * i] = sub.a;
        subs[2 * i + 1] = sub.b;
        k = nk;
    }
    key_io[0] = k.a;
    key_io[1] = k.b;
}

/* Play envs [env_lo, env_hi) of a global batch of `batch` envs to termination.
 * subs: (1 + 2*max_steps, 2) chain sub keys: [0] init, [1+2t] act keys of loop step t,
 * [2+2t] step keys.  Outputs are indexed from env_lo.  Returns the largest episode
 * length, or -1 if some env is still running after max_steps. */
int64_t orc_play(const uint32_t* subs, int64_t max_steps, int64_t batch, int64_t env_lo, int64_t env_hi,
                 int policy, int mode, uint8_t* final_boards, int32_t* lengths, int64_t* scores,
                 int32_t* first_actions /* may be NULL, else (n, 16) */) {
    int64_t longest = 0;
    int failed = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(max : longest) reduction(| : failed)
    for (int64_t e = env_lo; e < env_hi; ++e) {
        uint8_t b[16], mask[4], done = 0;
        key_t2 sub0 = {subs[0], subs[1]};
        init_one(split_at(sub0, (uint32_t)batch, (uint32_t)e, mode), mode, b, mask);
        int64_t score = 0, t = 0;
        for (; t < max_steps && !done; ++t) {
            key_t2 sa = {subs[2 * (1 + 2 * t)], subs[2 * (1 + 2 * t) + 1]};
            key_t2 ss = {subs[2 * (2 + 2 * t)], subs[2 * (2 + 2 * t) + 1]};
            int a = policy == POLICY_RANDOM
                        ? act_random_one(split_at(sa, (uint32_t)batch, (uint32_t)e, mode), mask, mode, NULL)
                        : act_drul_one(mask);
            if (first_actions && t < 16) first_actions[16 * (e - env_lo) + t] = a;
            float r = step_one(b, mask, &done, a, split_at(ss, (uint32_t)batch, (uint32_t)e, mode), mode);
            if (r > 0) score += (int64_t)r;
        }
        if (!done) failed = 1;
        memcpy(final_boards + 16 * (e - env_lo), b, 16);
        lengths[e - env_lo] = (int32_t)t;
        scores[e - env_lo] = score;
        if (t > longest) longest = t;
    }
    return failed ? -1 : longest;
}

/* src/ppo/data_loader.py:103-130 */
void orc_gae(const float* rewards, const float* values, const uint8_t* dones, int64_t n, double gamma,
             double lambda_gae, float* adv, float* ret) {
    const float g = (float)gamma, gl = (float)(gamma * lambda_gae);
    float last_gae = 0.0f, last_value = 0.0f;
    for (int64_t t = n - 1; t >= 0; --t) {
        if (dones[t]) {
            last_value = 0.0f;
            last_gae = 0.0f;
        }
        float delta = (rewards[t] + g * last_value) - values[t];
        last_gae = delta + gl * last_gae;
        adv[t] = last_gae;
        ret[t] = last_gae + values[t];
        last_value = values[t];
    }
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
