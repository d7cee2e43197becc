/* First-generation kernels, kept as measured baselines and as a second implementation to test the current kernels
 * against.  TEST INFRASTRUCTURE: they are compiled only into tests/legacy/libg2048_legacy.so (make -C
 * 2048-ppo-agent_b200/csrc legacy: the library's sources with -DG2048_LEGACY_KERNELS), which tests/conftest.py and
 * tools/ab_play.py load next to the product library; libg2048.so does not export them. */
#ifndef G2048_LEGACY_H_
#define G2048_LEGACY_H_
#include "../../include/g2048.h"
#ifdef __cplusplus
extern "C" {
#endif

/* First-generation play kernel (lanes park until six are free, per-step reward loop): identical
 * arguments and results; kept so that the current kernel can be A/B-timed against it. */
int g2048_play_v1(int policy, const uint32_t* d_subs, int64_t n_subs, int64_t batch_global, int64_t env_lo, int64_t n,
                  int rng_mode, uint64_t* d_work, uint64_t* d_final_boards, uint32_t* d_lengths, uint32_t* d_scores,
                  uint64_t* d_stats, void* stream);

/* First-generation observation kernel (one 16-byte chunk per thread, plain stores); arguments of g2048_expand_obs. */
int g2048_expand_obs_v1(const uint64_t* d_boards, int64_t n, int dtype, void* d_out, int64_t rows, int64_t n_cols,
                        void* stream);

/* First-generation GAE kernel (1 024-step tiles, every array staged in shared memory); arguments of g2048_gae_flat. */
int g2048_gae_flat_v1(const float* d_rewards, const float* d_values, const uint8_t* d_dones, int64_t n, double gamma,
                      double lambda_gae, float* d_adv, float* d_ret, void* d_scan_state, double* d_moments,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* G2048_LEGACY_H_ */
