// One-hot observation expansion rebuilt around the bulk copy engine (TMA, 1-D form): a warp keeps a
// ring of 1984-byte images in shared memory, flips the 16 ones of a board in place and lets one
// `cp.async.bulk` (SASS: UBLKCP) store the image -- ~25 warp instructions per board instead of ~150
// for the per-thread 16-byte-chunk kernel, which stays available as g2048_expand_obs_v1.
#include <cuda_bf16.h>

#include <cstdlib>

#include "g2048_common.cuh"
#include "g2048_env.cuh"
#include "g2048_tma.cuh"

namespace g2048 {

typedef unsigned long long u64;

// ------------------------------------------------------------------------------------------------
// one-hot observation expansion
// ------------------------------------------------------------------------------------------------
constexpr int OBS_THREADS = 256;
constexpr int OBS_WARPS = OBS_THREADS / 32;
constexpr int OBS_NBUF = 4;          // images in flight per warp
constexpr int OBS_IMAGE_BYTES = 1984;  // 496 f32 = 1 board, 992 bf16 = 2 boards, 1984 u8 = 4 boards
constexpr int OBS_SMEM_BYTES = OBS_WARPS * OBS_NBUF * OBS_IMAGE_BYTES;
// 0: 256-thread CTAs at every size (A/B timing of the one-warp launches of small batches)
#ifndef G2048_OBS_NARROW
#define G2048_OBS_NARROW 1
#endif

template <typename T> struct ObsOne;
template <> struct ObsOne<float> { static __device__ float one() { return 1.0f; } static __device__ float zero() { return 0.0f; } };
template <> struct ObsOne<__nv_bfloat16> {
    static __device__ __nv_bfloat16 one() { return __ushort_as_bfloat16((unsigned short)0x3F80); }
    static __device__ __nv_bfloat16 zero() { return __ushort_as_bfloat16((unsigned short)0); }
};
template <> struct ObsOne<uint8_t> { static __device__ uint8_t one() { return 1; } static __device__ uint8_t zero() { return 0; } };

// The per-sample scalars of a minibatch (g2048_gather_minibatch): gathered by the observation kernel itself, after its
// last bulk store has been issued and before it waits for the stores to drain, so the minibatch is ONE launch.
// (Tried and dropped: fetching them along with the boards inside the image loop, one scalar per idle lane, a round
// ahead of its use -- the scattered 4-byte stores between an image's flips and its fence.proxy.async slowed the loop
// from 243 to 385 us at 2^19 samples.  As a phase of their own the 2.6e6 random sector reads take ~40 us there.)
struct GatherScalars {
    const uint8_t* meta;
    const float *log_probs, *values, *adv, *ret;
    int64_t* o_actions;
    uchar4* o_masks;
    float *o_log_probs, *o_values, *o_adv, *o_ret;
    const unsigned long long* boards;  // stand-alone scalar kernel only (no observations: the batch carries bitboards)
    unsigned long long* o_boards;
    // sample-record form (g2048_pack_samples): every field of a sample sits in ONE 32-byte sector -- two 16-byte loads
    // per sample instead of one sector per source array; the pointers above are ignored when this is set
    const uint4* records;
};

// G2048SampleRecord (include/g2048.h): {board u64, meta u32, reward f32 | log_prob, value, advantage, return f32}
__device__ __forceinline__ void gather_record_at(const GatherScalars& g, const int64_t* __restrict__ idx, int64_t i) {
    const int64_t s = __ldg(&idx[i]);
    const uint4 h0 = __ldg(&g.records[2 * s]);
    const uint4 h1 = __ldg(&g.records[2 * s + 1]);
    const uint32_t mt = h0.z;
    if (g.o_actions) g.o_actions[i] = (int64_t)(mt & 3u);
    if (g.o_masks) g.o_masks[i] = make_uchar4((mt >> 2) & 1u, (mt >> 3) & 1u, (mt >> 4) & 1u, (mt >> 5) & 1u);
    if (g.o_log_probs) g.o_log_probs[i] = __uint_as_float(h1.x);
    if (g.o_values) g.o_values[i] = __uint_as_float(h1.y);
    if (g.o_adv) g.o_adv[i] = __uint_as_float(h1.z);
    if (g.o_ret) g.o_ret[i] = __uint_as_float(h1.w);
    if (g.o_boards) g.o_boards[i] = ((unsigned long long)h0.y << 32) | (unsigned long long)h0.x;
}

__device__ __forceinline__ void gather_scalars_at(const GatherScalars& g, const int64_t* __restrict__ idx, int64_t i) {
    if (g.records) return gather_record_at(g, idx, i);
    const int64_t s = __ldg(&idx[i]);
    const uint32_t mt = g.meta ? g.meta[s] : 0u;
    const float lp = (g.o_log_probs && g.log_probs) ? g.log_probs[s] : 0.0f;
    const float vl = (g.o_values && g.values) ? g.values[s] : 0.0f;
    const float ad = (g.o_adv && g.adv) ? g.adv[s] : 0.0f;
    const float rt = (g.o_ret && g.ret) ? g.ret[s] : 0.0f;
    if (g.o_actions) g.o_actions[i] = (int64_t)(mt & 3u);
    if (g.o_masks) g.o_masks[i] = make_uchar4((mt >> 2) & 1u, (mt >> 3) & 1u, (mt >> 4) & 1u, (mt >> 5) & 1u);
    if (g.o_log_probs && g.log_probs) g.o_log_probs[i] = lp;
    if (g.o_values && g.values) g.o_values[i] = vl;
    if (g.o_adv && g.adv) g.o_adv[i] = ad;
    if (g.o_ret && g.ret) g.o_ret[i] = rt;
    if (g.o_boards && g.boards) g.o_boards[i] = g.boards[s];
}

// AHEAD: fetch the boards of a whole round of images before the per-image loop.  With a gather (or the time-major
// index map) every board is its own memory round trip, and inside the loop those round trips run one after the
// other: 2^19-sample minibatch gather 4.2 -> 4.4 TB/s.  The contiguous expansion (boards share cache lines) loses
// 3 % to the extra registers, so it keeps the in-loop loads.
template <typename T, bool AHEAD, bool SCALARS = false>
__global__ void __launch_bounds__(OBS_THREADS)
expand_obs_tma_kernel(const u64* __restrict__ boards, int64_t n, T* __restrict__ out, int64_t rows, int64_t n_cols,
                      const int64_t* __restrict__ indices, const GatherScalars sc = GatherScalars{}, int board_stride = 1) {
    constexpr int G = OBS_IMAGE_BYTES / (496 * (int)sizeof(T));  // boards per image
    constexpr int CELLS = 16 * G;
    constexpr int PER_LANE = (CELLS + 31) / 32;
    extern __shared__ __align__(128) uint8_t s_img_raw[];  // [OBS_WARPS][OBS_NBUF][OBS_IMAGE_BYTES]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* ring = s_img_raw + (size_t)warp * OBS_NBUF * OBS_IMAGE_BYTES;
    // zero the ring once; afterwards only the ones are flipped
    for (int i = lane; i < OBS_NBUF * OBS_IMAGE_BYTES / 16; i += 32)
        reinterpret_cast<uint4*>(ring)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();

    const int64_t n_images = (n + G - 1) / G;
    const int64_t warps_total = (int64_t)gridDim.x * OBS_WARPS;
    int old_pos[OBS_NBUF][PER_LANE];
#pragma unroll
    for (int j = 0; j < OBS_NBUF; ++j)
#pragma unroll
        for (int p = 0; p < PER_LANE; ++p) old_pos[j][p] = -1;

    auto board_of = [&](int64_t src) -> u64 {
        if (rows > 0) {  // out is env-major (col, row); boards are time-major (row, col)
            const int64_t col = src / rows;
            src = (src - col * rows) * n_cols + col;
        }
        if (indices) src = __ldg(&indices[src]);  // gather: out[i] = onehot(boards[indices[i]])
        return __ldg(&boards[src * board_stride]);  // stride 4: the board of a 32-byte sample record
    };
    // Source positions and boards of one round of OBS_NBUF images (image j of the round starting at `base` is
    // base + j * warps_total).  With a gather a board is TWO dependent memory round trips (index, then board): the
    // positions are fetched two rounds ahead and the boards one round ahead, so that neither load is waited for while
    // it is in flight -- a warp used to spend ~1.5 us per round at the top of its loop (37 rounds per warp at 2^19
    // samples: a quarter of the kernel; ncu long_scoreboard 19 per issue).
    auto load_positions = [&](int64_t base, int64_t (&dst)[OBS_NBUF][PER_LANE]) {
#pragma unroll
        for (int j = 0; j < OBS_NBUF; ++j) {
            const int64_t first = (base + (int64_t)j * warps_total) * G;
#pragma unroll
            for (int p = 0; p < PER_LANE; ++p) {
                const int c = lane + 32 * p;
                int64_t src = first + (c >> 4);
                const bool in = c < CELLS && src < n;
                if (in && rows > 0) {  // out is env-major (col, row); boards are time-major (row, col)
                    const int64_t col = src / rows;
                    src = (src - col * rows) * n_cols + col;
                }
                if (in && indices) src = __ldg(&indices[src]);  // gather: out[i] = onehot(boards[indices[i]])
                dst[j][p] = in ? src : -1;
            }
        }
    };
    auto load_boards = [&](const int64_t (&pos)[OBS_NBUF][PER_LANE], u64 (&dst)[OBS_NBUF][PER_LANE]) {
#pragma unroll
        for (int j = 0; j < OBS_NBUF; ++j)
#pragma unroll
            for (int p = 0; p < PER_LANE; ++p)
                dst[j][p] = pos[j][p] >= 0 ? __ldg(&boards[pos[j][p] * board_stride]) : 0ull;  // stride 4: sample records
    };
    int64_t img = (int64_t)blockIdx.x * OBS_WARPS + warp;
    const int64_t round_stride = (int64_t)OBS_NBUF * warps_total;
    u64 cur_boards[OBS_NBUF][PER_LANE];     // dead code without AHEAD
    int64_t next_pos[OBS_NBUF][PER_LANE];   // positions of the round after the current one
    if (AHEAD) {
        if (img < n_images) {
            load_positions(img, next_pos);
            load_boards(next_pos, cur_boards);
        }
        if (img + round_stride < n_images) load_positions(img + round_stride, next_pos);
    }
    while (img < n_images) {
        u64 next_boards[OBS_NBUF][PER_LANE];
        if (AHEAD) {
            if (img + round_stride < n_images) load_boards(next_pos, next_boards);            // round r + 1: positions are here
            if (img + 2 * round_stride < n_images) load_positions(img + 2 * round_stride, next_pos);  // round r + 2
        }
#pragma unroll
        for (int j = 0; j < OBS_NBUF; ++j) {
            if (img >= n_images) break;  // warp-uniform
            T* buf = reinterpret_cast<T*>(ring + j * OBS_IMAGE_BYTES);
            // the store issued OBS_NBUF images ago read this buffer: wait until it is done with it
            if (lane == 0) bulk_wait_read<OBS_NBUF - 1>();
            __syncwarp();
            const int64_t first = img * G;
            const int in_image = (int)min((int64_t)G, n - first);
#pragma unroll
            for (int p = 0; p < PER_LANE; ++p) {
                const int c = lane + 32 * p;  // cell index inside the image
                if (c < CELLS) {
                    if (old_pos[j][p] >= 0) buf[old_pos[j][p]] = ObsOne<T>::zero();
                    const int g = c >> 4, cell = c & 15;
                    int pos = -1;
                    if (g < in_image) {
                        const u64 b = AHEAD ? cur_boards[j][p] : board_of(first + g);
                        pos = g * 496 + 31 * cell + (int)((b >> (4 * cell)) & 15ull);
                        buf[pos] = ObsOne<T>::one();
                    }
                    old_pos[j][p] = pos;
                }
            }
            fence_proxy_async();  // generic-proxy writes above -> visible to the bulk copy
            __syncwarp();
            if (lane == 0) {
                bulk_store(out + first * 496, buf, (uint32_t)(in_image * 496 * (int)sizeof(T)));
                bulk_commit();
            }
            img += warps_total;
        }
        if (AHEAD) {
#pragma unroll
            for (int j = 0; j < OBS_NBUF; ++j)
#pragma unroll
                for (int p = 0; p < PER_LANE; ++p) cur_boards[j][p] = next_boards[j][p];
        }
    }
    if (SCALARS) {  // the minibatch's per-sample scalars, while the last bulk stores drain
        for (int64_t i = (int64_t)blockIdx.x * OBS_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * OBS_THREADS)
            gather_scalars_at(sc, indices, i);
    }
    if (lane == 0) bulk_wait_all<0>();
}

}  // namespace g2048

using namespace g2048;

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

static int launch_expand_obs(const uint64_t* d_boards, int64_t n, int dtype, void* d_out, int64_t rows, int64_t n_cols,
                             const int64_t* d_indices, void* stream, const GatherScalars* scalars = nullptr,
                             int board_stride = 1) {
    G2048_REQUIRE(n >= 0 && rows >= 0 && (rows == 0 || (n_cols > 0 && rows * n_cols == n)), "expand_obs: shape");
    G2048_REQUIRE(!scalars || (d_indices && rows == 0), "expand_obs: scalars need an index list");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_boards && d_out && aligned16(d_out), "expand_obs: pointers (out must be 16-byte aligned)");
    const int sms = sm_count();
    if (sms <= 0) return fail_arg("expand_obs: no device");
    cudaStream_t st = (cudaStream_t)stream;
    static bool configured_on[64] = {false};
    bool* configured = device_once_flag(configured_on);
    if (!configured) return fail_arg("no CUDA device");
    if (!*configured) {
        int rc = G2048_OK;
#define G2048_OBS_SMEM(T, A) \
    if (!rc) rc = check_cuda(cudaFuncSetAttribute(expand_obs_tma_kernel<T, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, OBS_SMEM_BYTES), "expand_obs: smem attribute")
        G2048_OBS_SMEM(float, false); G2048_OBS_SMEM(float, true);
        G2048_OBS_SMEM(__nv_bfloat16, false); G2048_OBS_SMEM(__nv_bfloat16, true);
        G2048_OBS_SMEM(uint8_t, false); G2048_OBS_SMEM(uint8_t, true);
#undef G2048_OBS_SMEM
#define G2048_OBS_SMEM_S(T) \
    if (!rc) rc = check_cuda(cudaFuncSetAttribute(expand_obs_tma_kernel<T, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, OBS_SMEM_BYTES), "expand_obs: smem attribute")
        G2048_OBS_SMEM_S(float); G2048_OBS_SMEM_S(__nv_bfloat16); G2048_OBS_SMEM_S(uint8_t);
#undef G2048_OBS_SMEM_S
        if (rc) return rc;
        *configured = true;
    }
    const bool ahead = d_indices != nullptr || rows > 0;  // every board is its own round trip
    auto grid_for = [&](int64_t images) {
        const int64_t need = (images + OBS_WARPS - 1) / OBS_WARPS;
        const int64_t cap = (int64_t)sms * 3;  // 3 resident CTAs of 62 KiB per SM, persistent loop
        return (unsigned)(need < cap ? need : cap);
    };
#define G2048_OBS_LAUNCH(T, images)                                                                                     \
    do {                                                                                                                \
        if (scalars)                                                                                                    \
            expand_obs_tma_kernel<T, true, true><<<grid_for(images), OBS_THREADS, OBS_SMEM_BYTES, st>>>(                \
                (const u64*)d_boards, n, (T*)d_out, rows, n_cols, d_indices, *scalars, board_stride);                   \
        else if (ahead)                                                                                                 \
            expand_obs_tma_kernel<T, true><<<grid_for(images), OBS_THREADS, OBS_SMEM_BYTES, st>>>(                      \
                (const u64*)d_boards, n, (T*)d_out, rows, n_cols, d_indices);                                           \
        else                                                                                                            \
            expand_obs_tma_kernel<T, false><<<grid_for(images), OBS_THREADS, OBS_SMEM_BYTES, st>>>(                     \
                (const u64*)d_boards, n, (T*)d_out, rows, n_cols, d_indices);                                           \
    } while (0)
    switch (dtype) {
        case G2048_OBS_F32: G2048_OBS_LAUNCH(float, n); break;
        case G2048_OBS_BF16: G2048_OBS_LAUNCH(__nv_bfloat16, (n + 1) / 2); break;
        case G2048_OBS_BOOL: G2048_OBS_LAUNCH(uint8_t, (n + 3) / 4); break;
        default: return fail_arg("expand_obs: dtype");
    }
#undef G2048_OBS_LAUNCH
    G2048_CHECK_LAUNCH("expand_obs");
    return G2048_OK;
}

extern "C" int g2048_expand_obs(const uint64_t* d_boards, int64_t n, int dtype, void* d_out, int64_t rows,
                                int64_t n_cols, void* stream) {
    return launch_expand_obs(d_boards, n, dtype, d_out, rows, n_cols, nullptr, stream);
}

// out[i] = one-hot of d_boards[d_indices[i]], i < m (live-env compaction of the network-policy loop; the minibatch
// gather uses the same kernel)
extern "C" int g2048_expand_obs_gather(const uint64_t* d_boards, const int64_t* d_indices, int64_t m, int dtype, void* d_out,
                                       void* stream) {
    G2048_REQUIRE(d_indices || m == 0, "expand_obs_gather: indices");
    return launch_expand_obs(d_boards, m, dtype, d_out, 0, 0, d_indices, stream);
}

// ------------------------------------------------------------------------------------------------
// minibatch gather (replaces PPODataset.__getitem__ + DataLoader collation,
// src/ppo/data_loader.py:132-166,217-223, on the packed device buffer)
// ------------------------------------------------------------------------------------------------
namespace g2048 {
__global__ void __launch_bounds__(256)
gather_scalars_kernel(const int64_t* __restrict__ idx, int64_t m, const uint8_t* __restrict__ meta,
                      const float* __restrict__ log_probs, const float* __restrict__ values,
                      const float* __restrict__ adv, const float* __restrict__ ret, int64_t* __restrict__ o_actions,
                      uchar4* __restrict__ o_masks, float* __restrict__ o_log_probs, float* __restrict__ o_values,
                      float* __restrict__ o_adv, float* __restrict__ o_ret, const unsigned long long* __restrict__ boards,
                      unsigned long long* __restrict__ o_boards) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    gather_scalars_at(GatherScalars{meta, log_probs, values, adv, ret, o_actions, o_masks, o_log_probs, o_values, o_adv, o_ret,
                                    boards, o_boards, nullptr},
                      idx, i);
}
__global__ void __launch_bounds__(256) gather_records_kernel(const int64_t* __restrict__ idx, int64_t m, const GatherScalars sc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) gather_record_at(sc, idx, i);
}
}  // namespace g2048

extern "C" int g2048_gather_minibatch(const int64_t* d_indices, int64_t m, const uint64_t* d_boards,
                                      const uint8_t* d_meta, const float* d_log_probs, const float* d_values,
                                      const float* d_adv, const float* d_ret, int obs_dtype, void* d_obs,
                                      int64_t* d_actions, uint8_t* d_masks, float* d_old_log_probs, float* d_old_values,
                                      float* d_out_adv, float* d_out_ret, uint64_t* d_out_boards, void* stream) {
    G2048_REQUIRE(m >= 0, "gather_minibatch: m");
    if (m == 0) return G2048_OK;
    G2048_REQUIRE(d_indices, "gather_minibatch: indices");
    G2048_REQUIRE(d_boards || !(d_obs || d_out_boards), "gather_minibatch: boards");
    if (d_obs) {  // one launch: the observation kernel gathers the scalars while its last stores drain
        const g2048::GatherScalars sc{d_meta, d_log_probs, d_values, d_adv, d_ret, d_actions, (uchar4*)d_masks,
                                      d_old_log_probs, d_old_values, d_out_adv, d_out_ret,
                                      (const unsigned long long*)d_boards, (unsigned long long*)d_out_boards, nullptr};
        return launch_expand_obs(d_boards, m, obs_dtype, d_obs, 0, 0, d_indices, stream, &sc);
    }
    g2048::gather_scalars_kernel<<<blocks_for(m, 256), 256, 0, (cudaStream_t)stream>>>(
        d_indices, m, d_meta, d_log_probs, d_values, d_adv, d_ret, d_actions, (uchar4*)d_masks, d_old_log_probs,
        d_old_values, d_out_adv, d_out_ret, (const unsigned long long*)d_boards, (unsigned long long*)d_out_boards);
    G2048_CHECK_LAUNCH("gather_minibatch");
    return G2048_OK;
}

// ------------------------------------------------------------------------------------------------
// network-policy loop step fused with the NEXT step's observation (one launch per loop step)
// ------------------------------------------------------------------------------------------------
// g2048_policy_step (g2048_policy.cu) followed by g2048_expand_obs of the stepped boards, in one kernel: a warp owns 32
// envs -- each lane samples its action from the network's logits, steps its env and writes its record exactly as
// policy_step_kernel does -- and then expands the 32 NEW boards (handed from lane to lane by shuffles, they never
// leave the registers) into the observation tensor of the next forward pass through the same shared-memory image ring
// and bulk stores as expand_obs_tma_kernel.  Removes one launch per step and the 8 + 9 bytes per env of the state's
// round trip through HBM between the two kernels; the stores of a tile overlap the next tile's sampling.
namespace g2048 {

struct PolicyStepObsArgs {
    u64* boards;
    uint8_t* status;
    const float4* logits;
    const float* values;
    int use_mask, sample, auto_reset;
    const uint32_t* subs;        // act sub key of step t at words [4t, 4t+2), step sub key at [4t+2, 4t+4)
    int32_t* step_index;         // t (device memory; NULL: t = 0)
    int advance_step;            // the last CTA to finish adds 1 to *step_index (graph replay: no separate counter kernel)
    uint32_t batch_global, env_lo;
    int64_t n;
    u64* rec_boards;             // record bases; slot t * n is written
    uint8_t* rec_meta;
    float *rec_rewards, *rec_log_probs, *rec_values;
    int32_t* actions_out;
    unsigned long long* counters;  // [0] += envs that terminated on this step, [1] += reward > 0 sum (may be NULL)
};

__device__ unsigned g_step_ticket = 0u;  // CTAs of the running policy_step_obs launch that have read the step number and finished

template <int MODE, typename T>
__global__ void __launch_bounds__(OBS_THREADS)
policy_step_obs_kernel(const PolicyStepObsArgs a, T* __restrict__ obs) {
    constexpr int G = OBS_IMAGE_BYTES / (496 * (int)sizeof(T));  // boards per image
    constexpr int CELLS = 16 * G;
    constexpr int PER_LANE = (CELLS + 31) / 32;
    constexpr int IMAGES = 32 / G;  // images per 32-env tile
    extern __shared__ __align__(128) uint8_t s_img_raw[];  // [OBS_WARPS][OBS_NBUF][OBS_IMAGE_BYTES]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* ring = s_img_raw + (size_t)warp * OBS_NBUF * OBS_IMAGE_BYTES;
    for (int i = lane; i < OBS_NBUF * OBS_IMAGE_BYTES / 16; i += 32)
        reinterpret_cast<uint4*>(ring)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    int old_pos[OBS_NBUF][PER_LANE];
#pragma unroll
    for (int j = 0; j < OBS_NBUF; ++j)
#pragma unroll
        for (int p = 0; p < PER_LANE; ++p) old_pos[j][p] = -1;

    const int64_t t = a.step_index ? (int64_t)*a.step_index : 0;
    const uint32_t* sub_act = a.subs + 4 * t;
    const uint32_t* sub_step = sub_act + 2;
    const int64_t at = t * a.n;
    const int64_t n_tiles = (a.n + 31) / 32;
    unsigned long long done_new = 0, reward_sum = 0;

    // warp w of CTA b takes tiles w * gridDim + b, + OBS_WARPS * gridDim, ...: every CTA (hence every SM) gets its share
    for (int64_t tile = (int64_t)warp * gridDim.x + blockIdx.x; tile < n_tiles; tile += (int64_t)OBS_WARPS * gridDim.x) {
        const int64_t i = tile * 32 + lane;
        const bool valid = i < a.n;
        u64 nb = 0ull;
        if (valid) {  // ---- policy_step_kernel's body for env i --------------------------------------------------
            EnvState s{a.boards[i], a.status[i]};
            const bool was_done = (s.status & G2048_STATUS_DONE) != 0u;
            if (a.auto_reset && was_done) s.status &= ~(uint32_t)G2048_STATUS_DONE;
            const uint32_t lm = s.status & G2048_STATUS_MASK;
            const Logits4 l = prepare_logits(a.logits[i], lm, a.use_mask != 0);
            int act;
            if (a.sample) {
                act = sample_categorical<MODE>(env_key<MODE>(sub_act, a.batch_global, a.env_lo, i), l);
            } else {
                act = argmax4(l);
            }
            const float picked = act == 0 ? l.v[0] : (act == 1 ? l.v[1] : (act == 2 ? l.v[2] : l.v[3]));
            const float lp = picked - log_sum_exp4(l);
            const u64 pre = s.board;
            const Key step_key = env_key<MODE>(sub_step, a.batch_global, a.env_lo, i);
            float r;
            if (a.auto_reset) {
                Key k1, k2;
                split2<MODE>(step_key, k1, k2);
                r = env_step<MODE>(s, act, k1);
                if (s.status & G2048_STATUS_DONE) {
                    const EnvState fresh = env_init<MODE>(k2);
                    s.board = fresh.board;
                    s.status = fresh.status | G2048_STATUS_DONE | (s.status & G2048_STATUS_OVERFLOW);
                }
            } else {
                r = env_step<MODE>(s, act, step_key);
            }
            a.boards[i] = s.board;
            a.status[i] = (uint8_t)s.status;
            const bool done = (s.status & G2048_STATUS_DONE) != 0u;
            if (a.rec_boards) a.rec_boards[at + i] = pre;
            if (a.rec_meta) a.rec_meta[at + i] = (uint8_t)((uint32_t)act | (lm << 2) | (done ? 0x40u : 0u));
            if (a.rec_rewards) a.rec_rewards[at + i] = r;
            if (a.rec_log_probs) a.rec_log_probs[at + i] = lp;
            if (a.rec_values && a.values) a.rec_values[at + i] = a.values[i];
            if (a.actions_out) a.actions_out[i] = act;
            if (done && (a.auto_reset || !was_done)) done_new += 1;
            if (r > 0.0f) reward_sum += (unsigned long long)r;
            nb = s.board;
        }
        // ---- the observation of the next forward pass: the tile's new boards, image by image ------------------------
        const int in_tile = (int)min((int64_t)32, a.n - tile * 32);
#pragma unroll
        for (int j0 = 0; j0 < IMAGES; j0 += OBS_NBUF) {
#pragma unroll
            for (int jj = 0; jj < OBS_NBUF; ++jj) {
                const int j = j0 + jj;                 // image of the tile; buffer jj of the ring (IMAGES % OBS_NBUF == 0)
                const int in_image = min(G, in_tile - j * G);  // warp-uniform
                T* buf = reinterpret_cast<T*>(ring + jj * OBS_IMAGE_BYTES);
                if (in_image > 0) {
                    if (lane == 0) bulk_wait_read<OBS_NBUF - 1>();  // the store issued OBS_NBUF images ago is done with buf
                    __syncwarp();
                }
#pragma unroll
                for (int p = 0; p < PER_LANE; ++p) {
                    const int c = lane + 32 * p;  // cell index inside the image
                    const int g = c >> 4, cell = c & 15;
                    const u64 b = __shfl_sync(0xFFFFFFFFu, nb, (j * G + g) & 31);  // every lane takes part
                    if (in_image > 0 && c < CELLS) {
                        if (old_pos[jj][p] >= 0) buf[old_pos[jj][p]] = ObsOne<T>::zero();
                        int pos = -1;
                        if (g < in_image) {
                            pos = g * 496 + 31 * cell + (int)((b >> (4 * cell)) & 15ull);
                            buf[pos] = ObsOne<T>::one();
                        }
                        old_pos[jj][p] = pos;
                    }
                }
                if (in_image > 0) {
                    fence_proxy_async();  // generic-proxy writes above -> visible to the bulk copy
                    __syncwarp();
                    if (lane == 0) {
                        bulk_store(obs + (tile * 32 + (int64_t)j * G) * 496, buf, (uint32_t)(in_image * 496 * (int)sizeof(T)));
                        bulk_commit();
                    }
                }
            }
        }
    }
    if (a.counters) {
        for (int off = 16; off > 0; off >>= 1) {
            done_new += __shfl_down_sync(0xFFFFFFFFu, done_new, off);
            reward_sum += __shfl_down_sync(0xFFFFFFFFu, reward_sum, off);
        }
        if (lane == 0) {
            if (done_new) atomicAdd(&a.counters[0], done_new);
            if (reward_sum) atomicAdd(&a.counters[1], reward_sum);
        }
    }
    if (lane == 0) bulk_wait_all<0>();
    if (a.step_index && a.advance_step) {  // every CTA read t at its start; the last one to get here moves it on
        __syncthreads();
        if (threadIdx.x == 0) {
            race_jitter();
            __threadfence();
            if (atomicAdd(&g_step_ticket, 1u) == gridDim.x - 1u) {
                g_step_ticket = 0u;
                *a.step_index = (int32_t)t + 1;
            }
        }
    }
}

}  // namespace g2048

extern "C" int g2048_policy_step_obs(uint64_t* d_boards, uint8_t* d_status, const float* d_logits, const float* d_values,
                                     int use_mask, int sample, int auto_reset, const uint32_t* d_subs,
                                     int32_t* d_step_index, int advance_step, int64_t batch_global, int64_t env_lo, int64_t n,
                                     int rng_mode, uint64_t* d_rec_boards, uint8_t* d_rec_meta, float* d_rec_rewards,
                                     float* d_rec_log_probs, float* d_rec_values, int32_t* d_actions_out, int obs_dtype,
                                     void* d_obs_next, uint64_t* d_counters, void* stream) {
    G2048_REQUIRE(rng_mode == G2048_RNG_ORIGINAL || rng_mode == G2048_RNG_PARTITIONABLE, "policy_step_obs: rng_mode");
    G2048_REQUIRE(batch_global > 0 && batch_global <= 0x7FFFFFFFll && env_lo >= 0 && n >= 0 && env_lo + n <= batch_global,
                  "policy_step_obs: batch");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_boards && d_status && d_logits && d_subs && d_obs_next, "policy_step_obs: pointers");
    G2048_REQUIRE(aligned16(d_logits) && aligned16(d_obs_next), "policy_step_obs: logits and observations must be 16-byte aligned");
    const int sms = sm_count();
    if (sms <= 0) return fail_arg("policy_step_obs: no device");
    static bool configured_on[64] = {false};
    bool* configured = device_once_flag(configured_on);
    if (!configured) return fail_arg("no CUDA device");
    if (!*configured) {
        int rc = G2048_OK;
#define G2048_PSO_SMEM(M, T) \
    if (!rc) rc = check_cuda(cudaFuncSetAttribute(g2048::policy_step_obs_kernel<M, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, OBS_SMEM_BYTES), "policy_step_obs: smem attribute")
        G2048_PSO_SMEM(0, float); G2048_PSO_SMEM(1, float); G2048_PSO_SMEM(0, __nv_bfloat16); G2048_PSO_SMEM(1, __nv_bfloat16);
        G2048_PSO_SMEM(0, uint8_t); G2048_PSO_SMEM(1, uint8_t);
#undef G2048_PSO_SMEM
        if (rc) return rc;
        *configured = true;
    }
    const g2048::PolicyStepObsArgs args{(g2048::u64*)d_boards, d_status, (const float4*)d_logits, d_values, use_mask, sample,
                                        auto_reset, d_subs, d_step_index, advance_step, (uint32_t)batch_global, (uint32_t)env_lo, n,
                                        (g2048::u64*)d_rec_boards, d_rec_meta, d_rec_rewards, d_rec_log_probs, d_rec_values,
                                        d_actions_out, (unsigned long long*)d_counters};
    const int64_t n_tiles = (n + 31) / 32;
    const int64_t need = (n_tiles + OBS_WARPS - 1) / OBS_WARPS;
    static const int ctas_per_sm = [] { const char* e = getenv("G2048_PSO_CTAS_PER_SM"); const int v = e ? atoi(e) : 0; return v >= 1 && v <= 3 ? v : 3; }();
    const int64_t cap = (int64_t)sms * ctas_per_sm;  // up to 3 resident CTAs of 62 KiB per SM
    // a whole number of CTAs per SM once there is more than one CTA's worth of tiles per SM: 65 536 envs are 256 CTAs of
    // eight tiles -- two on 108 SMs, one on 40 -- or 296 CTAs of seven (tiles go round-robin over the CTAs): 28.9 -> 28.4 us
    unsigned grid = (unsigned)(need < cap ? need : cap);
    if (need > sms && need < cap) grid = (unsigned)(((need + sms - 1) / sms) * sms);
    static const int grid_override = [] { const char* e = getenv("G2048_PSO_GRID"); return e ? atoi(e) : 0; }();  // probing only
    if (grid_override > 0) grid = (unsigned)(grid_override < n_tiles ? grid_override : n_tiles);
    // Small batches (fewer tiles than the GPU has schedulers): one warp per CTA, one CTA per tile -- a step of 512 envs
    // is 16 warps, and as two 8-warp CTAs they share two SMs' schedulers and copy engines; spread out, each has its own.
    // (The kernel's tile loop visits tile blockIdx + warp * gridDim first: with one warp and gridDim = tiles that is all.)
    unsigned threads = OBS_THREADS;
    int smem = OBS_SMEM_BYTES;
    if (G2048_OBS_NARROW && grid_override <= 0 && n_tiles <= (int64_t)sms * 4) {
        threads = 32;
        smem = OBS_NBUF * OBS_IMAGE_BYTES;
        grid = (unsigned)n_tiles;
    }
    cudaStream_t st = (cudaStream_t)stream;
#define G2048_PSO_LAUNCH(T)                                                                                            \
    do {                                                                                                               \
        if (rng_mode == G2048_RNG_PARTITIONABLE)                                                                       \
            g2048::policy_step_obs_kernel<1, T><<<grid, threads, smem, st>>>(args, (T*)d_obs_next);                    \
        else                                                                                                           \
            g2048::policy_step_obs_kernel<0, T><<<grid, threads, smem, st>>>(args, (T*)d_obs_next);                    \
    } while (0)
    switch (obs_dtype) {
        case G2048_OBS_F32: G2048_PSO_LAUNCH(float); break;
        case G2048_OBS_BF16: G2048_PSO_LAUNCH(__nv_bfloat16); break;
        case G2048_OBS_BOOL: G2048_PSO_LAUNCH(uint8_t); break;
        default: return fail_arg("policy_step_obs: dtype");
    }
#undef G2048_PSO_LAUNCH
    G2048_CHECK_LAUNCH("policy_step_obs");
    return G2048_OK;
}

// ------------------------------------------------------------------------------------------------
// sample records: the training view of the flat buffer, one 32-byte sector per sample
// ------------------------------------------------------------------------------------------------
namespace g2048 {

// x -> (x - mean) / (std_unbiased + 1e-8) exactly as normalize_kernel (g2048_gae.cu) computes it
struct NormParams {
    float mean, denom;
};
__device__ __forceinline__ NormParams norm_params(const double* __restrict__ moments, int which) {
    const double cnt = moments[0];
    const double sum = moments[which], sumsq = moments[which + 1];
    const double mean_d = sum / cnt;
    const double var_d = (sumsq - sum * mean_d) / (cnt - 1.0);
    return NormParams{(float)mean_d, (float)sqrt(var_d > 0.0 ? var_d : 0.0) + 1e-8f};
}

__global__ void __launch_bounds__(256)
pack_samples_kernel(const u64* __restrict__ boards, const uint8_t* __restrict__ meta, const float* __restrict__ rewards,
                    const float* __restrict__ log_probs, const float* __restrict__ values, const float* __restrict__ adv,
                    const float* __restrict__ ret, int64_t n, const double* __restrict__ moments, uint4* __restrict__ records) {
    NormParams na{0.0f, 1.0f}, nr{0.0f, 1.0f};
    if (moments) {
        na = norm_params(moments, 1);
        nr = norm_params(moments, 3);
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // four steps per thread and iteration, every load issued before the first store: with one step per iteration the
    // 29 bytes a thread had in flight left the SMs at half of the bandwidth-delay product (ncu: long_scoreboard 25 per issue)
    constexpr int U = 4;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += U * stride) {
        u64 b[U];
        uint32_t mt[U];
        float rw[U], lp[U], vl[U], a[U], r[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int64_t i = i0 + k * stride;
            const bool in = i < n;
            b[k] = in ? boards[i] : 0ull;
            mt[k] = in ? (uint32_t)meta[i] : 0u;
            rw[k] = (in && rewards) ? rewards[i] : 0.0f;
            lp[k] = (in && log_probs) ? log_probs[i] : 0.0f;
            vl[k] = (in && values) ? values[i] : 0.0f;
            a[k] = (in && adv) ? adv[i] : 0.0f;
            r[k] = (in && ret) ? ret[i] : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int64_t i = i0 + k * stride;
            if (i < n) {
                if (moments) {
                    a[k] = (a[k] - na.mean) / na.denom;
                    r[k] = (r[k] - nr.mean) / nr.denom;
                }
                records[2 * i] = make_uint4((uint32_t)b[k], (uint32_t)(b[k] >> 32), mt[k], __float_as_uint(rw[k]));
                records[2 * i + 1] = make_uint4(__float_as_uint(lp[k]), __float_as_uint(vl[k]), __float_as_uint(a[k]), __float_as_uint(r[k]));
            }
        }
    }
}

}  // namespace g2048

namespace g2048 {

// Minibatch gather from sample records, tile form: a warp owns 32 CONSECUTIVE samples.  Lane l fetches the index and
// then the whole 32-byte record of sample l (all 32 records of the tile are in flight at once, and the next tile's
// indices and records are requested before this tile's images are written), writes its sample's scalars -- the warp's
// stores are coalesced -- and the 32 boards go from lane to lane by shuffles into the image ring, as in
// policy_step_obs_kernel.  Every record is touched ONCE.  (The round-robin form above, reading records: boards four
// per round, scalars in a phase of their own at the end -- by then 1 GB of observations has passed through L2 and
// every record comes from DRAM a second time, as a 128-byte line: ncu 126 MB read for 2^19 samples.)
template <typename T>
__global__ void __launch_bounds__(OBS_THREADS)
gather_samples_tile_kernel(const int64_t* __restrict__ indices, int64_t m, const uint4* __restrict__ records,
                           T* __restrict__ obs, const GatherScalars g) {
    constexpr int G = OBS_IMAGE_BYTES / (496 * (int)sizeof(T));  // boards per image
    constexpr int CELLS = 16 * G;
    constexpr int PER_LANE = (CELLS + 31) / 32;
    constexpr int IMAGES = 32 / G;  // images per 32-sample tile
    extern __shared__ __align__(128) uint8_t s_img_raw[];  // [OBS_WARPS][OBS_NBUF][OBS_IMAGE_BYTES]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* ring = s_img_raw + (size_t)warp * OBS_NBUF * OBS_IMAGE_BYTES;
    for (int i = lane; i < OBS_NBUF * OBS_IMAGE_BYTES / 16; i += 32)
        reinterpret_cast<uint4*>(ring)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    int old_pos[OBS_NBUF][PER_LANE];
#pragma unroll
    for (int j = 0; j < OBS_NBUF; ++j)
#pragma unroll
        for (int p = 0; p < PER_LANE; ++p) old_pos[j][p] = -1;

    const int64_t n_tiles = (m + 31) / 32;
    const int64_t stride = (int64_t)OBS_WARPS * gridDim.x;
    // warp w of CTA b takes tiles w * gridDim + b, + OBS_WARPS * gridDim, ...: every CTA (hence every SM) gets its share
    int64_t tile = (int64_t)warp * gridDim.x + blockIdx.x;
    auto fetch = [&](int64_t t, uint4& h0, uint4& h1) {
        const int64_t i = t * 32 + lane;
        if (t < n_tiles && i < m) {
            const int64_t s = __ldg(&indices[i]);
            h0 = __ldg(&records[2 * s]);
            h1 = __ldg(&records[2 * s + 1]);
        } else {
            h0 = make_uint4(0u, 0u, 0u, 0u);
            h1 = h0;
        }
    };
    uint4 h0, h1;
    fetch(tile, h0, h1);
    for (; tile < n_tiles; tile += stride) {
        uint4 n0, n1;
        fetch(tile + stride, n0, n1);  // the next tile's records, in flight while this tile's images are written
        const int64_t i = tile * 32 + lane;
        if (i < m) {  // my sample's scalars (G2048SampleRecord): coalesced over the warp
            const uint32_t mt = h0.z;
            if (g.o_actions) g.o_actions[i] = (int64_t)(mt & 3u);
            if (g.o_masks) g.o_masks[i] = make_uchar4((mt >> 2) & 1u, (mt >> 3) & 1u, (mt >> 4) & 1u, (mt >> 5) & 1u);
            if (g.o_log_probs) g.o_log_probs[i] = __uint_as_float(h1.x);
            if (g.o_values) g.o_values[i] = __uint_as_float(h1.y);
            if (g.o_adv) g.o_adv[i] = __uint_as_float(h1.z);
            if (g.o_ret) g.o_ret[i] = __uint_as_float(h1.w);
            if (g.o_boards) g.o_boards[i] = ((unsigned long long)h0.y << 32) | (unsigned long long)h0.x;
        }
        const u64 nb = ((u64)h0.y << 32) | (u64)h0.x;
        const int in_tile = (int)min((int64_t)32, m - tile * 32);
#pragma unroll
        for (int j0 = 0; j0 < IMAGES; j0 += OBS_NBUF) {
#pragma unroll
            for (int jj = 0; jj < OBS_NBUF; ++jj) {
                const int j = j0 + jj;                 // image of the tile; buffer jj of the ring (IMAGES % OBS_NBUF == 0)
                const int in_image = min(G, in_tile - j * G);  // warp-uniform
                T* buf = reinterpret_cast<T*>(ring + jj * OBS_IMAGE_BYTES);
                if (in_image > 0) {
                    if (lane == 0) bulk_wait_read<OBS_NBUF - 1>();  // the store issued OBS_NBUF images ago is done with buf
                    __syncwarp();
                }
#pragma unroll
                for (int p = 0; p < PER_LANE; ++p) {
                    const int c = lane + 32 * p;  // cell index inside the image
                    const int gg = c >> 4, cell = c & 15;
                    const u64 b = __shfl_sync(0xFFFFFFFFu, nb, (j * G + gg) & 31);  // every lane takes part
                    if (in_image > 0 && c < CELLS) {
                        if (old_pos[jj][p] >= 0) buf[old_pos[jj][p]] = ObsOne<T>::zero();
                        int pos = -1;
                        if (gg < in_image) {
                            pos = gg * 496 + 31 * cell + (int)((b >> (4 * cell)) & 15ull);
                            buf[pos] = ObsOne<T>::one();
                        }
                        old_pos[jj][p] = pos;
                    }
                }
                if (in_image > 0) {
                    fence_proxy_async();  // generic-proxy writes above -> visible to the bulk copy
                    __syncwarp();
                    if (lane == 0) {
                        bulk_store(obs + (tile * 32 + (int64_t)j * G) * 496, buf, (uint32_t)(in_image * 496 * (int)sizeof(T)));
                        bulk_commit();
                    }
                }
            }
        }
        h0 = n0;
        h1 = n1;
    }
    if (lane == 0) bulk_wait_all<0>();
}

}  // namespace g2048

extern "C" int g2048_pack_samples(const uint64_t* d_boards, const uint8_t* d_meta, const float* d_rewards,
                                  const float* d_log_probs, const float* d_values, const float* d_adv, const float* d_ret,
                                  int64_t n, const double* d_moments, G2048SampleRecord* d_records, void* stream) {
    G2048_REQUIRE(n >= 0, "pack_samples: n");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_boards && d_meta && d_records && aligned16(d_records), "pack_samples: pointers (records 16-byte aligned)");
    const int sms = sm_count();
    if (sms <= 0) return fail_arg("pack_samples: no device");
    int64_t g = (n + 255) / 256;
    const int64_t cap = (int64_t)sms * 8 * 4;
    if (g > cap) g = cap;
    g2048::pack_samples_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(
        (const g2048::u64*)d_boards, d_meta, d_rewards, d_log_probs, d_values, d_adv, d_ret, n, d_moments, (uint4*)d_records);
    G2048_CHECK_LAUNCH("pack_samples");
    return G2048_OK;
}

extern "C" int g2048_gather_samples(const int64_t* d_indices, int64_t m, const G2048SampleRecord* d_records, int obs_dtype,
                                    void* d_obs, int64_t* d_actions, uint8_t* d_masks, float* d_old_log_probs,
                                    float* d_old_values, float* d_out_adv, float* d_out_ret, uint64_t* d_out_boards,
                                    void* stream) {
    G2048_REQUIRE(m >= 0, "gather_samples: m");
    if (m == 0) return G2048_OK;
    G2048_REQUIRE(d_indices && d_records && aligned16(d_records), "gather_samples: pointers (records 16-byte aligned)");
    const g2048::GatherScalars sc{nullptr, nullptr, nullptr, nullptr, nullptr, d_actions, (uchar4*)d_masks, d_old_log_probs,
                                  d_old_values, d_out_adv, d_out_ret, nullptr, (unsigned long long*)d_out_boards,
                                  (const uint4*)d_records};
    if (d_obs) {
        static const bool round_robin = [] { const char* e = getenv("G2048_GATHER_ROUND_ROBIN"); return e && e[0] == '1'; }();
        if (round_robin)  // round 2's first form (A/B): the observation kernel reads the board out of the record (stride 4
                          // words) and gathers the scalars in a phase at its end
            return launch_expand_obs((const uint64_t*)d_records, m, obs_dtype, d_obs, 0, 0, d_indices, stream, &sc, 4);
        G2048_REQUIRE(aligned16(d_obs), "gather_samples: observations must be 16-byte aligned");
        const int sms = sm_count();
        if (sms <= 0) return fail_arg("gather_samples: no device");
        static bool configured_on[64] = {false};
        bool* configured = device_once_flag(configured_on);
        if (!configured) return fail_arg("no CUDA device");
        if (!*configured) {
            int rc = check_cuda(cudaFuncSetAttribute(g2048::gather_samples_tile_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, OBS_SMEM_BYTES), "gather_samples: smem attribute");
            if (!rc) rc = check_cuda(cudaFuncSetAttribute(g2048::gather_samples_tile_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, OBS_SMEM_BYTES), "gather_samples: smem attribute");
            if (!rc) rc = check_cuda(cudaFuncSetAttribute(g2048::gather_samples_tile_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, OBS_SMEM_BYTES), "gather_samples: smem attribute");
            if (rc) return rc;
            *configured = true;
        }
        const int64_t n_tiles = (m + 31) / 32;
        const int64_t need = (n_tiles + OBS_WARPS - 1) / OBS_WARPS;
        const int64_t cap = (int64_t)sms * 3;  // 3 resident CTAs of 62 KiB per SM
        unsigned grid = (unsigned)(need < cap ? need : cap);
        if (need > sms && need < cap) grid = (unsigned)(((need + sms - 1) / sms) * sms);  // whole CTAs per SM, as in g2048_policy_step_obs
        unsigned threads = OBS_THREADS;
        int smem = OBS_SMEM_BYTES;
        if (G2048_OBS_NARROW && n_tiles <= (int64_t)sms * 4) {  // a minibatch-sized gather: one warp per CTA, one CTA per tile (see g2048_policy_step_obs)
            threads = 32;
            smem = OBS_NBUF * OBS_IMAGE_BYTES;
            grid = (unsigned)n_tiles;
        }
        cudaStream_t st = (cudaStream_t)stream;
        switch (obs_dtype) {
            case G2048_OBS_F32:
                g2048::gather_samples_tile_kernel<float><<<grid, threads, smem, st>>>(d_indices, m, (const uint4*)d_records, (float*)d_obs, sc);
                break;
            case G2048_OBS_BF16:
                g2048::gather_samples_tile_kernel<__nv_bfloat16><<<grid, threads, smem, st>>>(d_indices, m, (const uint4*)d_records, (__nv_bfloat16*)d_obs, sc);
                break;
            case G2048_OBS_BOOL:
                g2048::gather_samples_tile_kernel<uint8_t><<<grid, threads, smem, st>>>(d_indices, m, (const uint4*)d_records, (uint8_t*)d_obs, sc);
                break;
            default: return fail_arg("gather_samples: dtype");
        }
        G2048_CHECK_LAUNCH("gather_samples");
        return G2048_OK;
    }
    g2048::gather_records_kernel<<<blocks_for(m, 256), 256, 0, (cudaStream_t)stream>>>(d_indices, m, sc);
    G2048_CHECK_LAUNCH("gather_samples");
    return G2048_OK;
}
