// HBM-bound record kernels: one-hot observation expansion, (T,B) -> (B,T) unpacking, the ragged
// env-major compaction of RolloutBuffer.store_batch, and the small reductions around them.
// Every kernel streams: coalesced reads, 128-bit coalesced stores, no re-reads.
#include <cuda_bf16.h>

#include <algorithm>

#include "g2048_common.cuh"
#include "g2048_rng.cuh"

namespace g2048 {

typedef unsigned long long u64;

// ------------------------------------------------------------------------------------------------
// one-hot observation: (n,16,31) of T.  One thread writes one 16-byte chunk; a chunk of V elements
// (V <= 16 < 31) overlaps at most two cells, so it holds at most two ones.
// Algorithmic bytes per board: 8 read + 496*sizeof(T) written (f32: 1992 B).
// ------------------------------------------------------------------------------------------------
#ifdef G2048_LEGACY_KERNELS  // first-generation observation kernel: built only into the tests' libg2048_legacy.so
template <typename T> struct OneHotChunk;

template <> struct OneHotChunk<float> {
    static constexpr int V = 4;
    __device__ static uint4 make(int rel0, int rel1) {
        const uint32_t one = 0x3F800000u;
        uint4 o;
        o.x = (rel0 == 0 || rel1 == 0) ? one : 0u;
        o.y = (rel0 == 1 || rel1 == 1) ? one : 0u;
        o.z = (rel0 == 2 || rel1 == 2) ? one : 0u;
        o.w = (rel0 == 3 || rel1 == 3) ? one : 0u;
        return o;
    }
};

template <> struct OneHotChunk<__nv_bfloat16> {
    static constexpr int V = 8;
    __device__ static uint32_t word(int rel, int w) {
        return ((rel >> 1) == w) ? (0x3F80u << (16 * (rel & 1))) : 0u;  // rel < 0 never matches w >= 0
    }
    __device__ static uint4 make(int rel0, int rel1) {
        uint4 o;
        o.x = word(rel0, 0) | word(rel1, 0);
        o.y = word(rel0, 1) | word(rel1, 1);
        o.z = word(rel0, 2) | word(rel1, 2);
        o.w = word(rel0, 3) | word(rel1, 3);
        return o;
    }
};

template <> struct OneHotChunk<uint8_t> {
    static constexpr int V = 16;
    __device__ static uint32_t word(int rel, int w) { return ((rel >> 2) == w) ? (1u << (8 * (rel & 3))) : 0u; }
    __device__ static uint4 make(int rel0, int rel1) {
        uint4 o;
        o.x = word(rel0, 0) | word(rel1, 0);
        o.y = word(rel0, 1) | word(rel1, 1);
        o.z = word(rel0, 2) | word(rel1, 2);
        o.w = word(rel0, 3) | word(rel1, 3);
        return o;
    }
};

template <typename T>
__global__ void __launch_bounds__(256)
expand_obs_kernel(const u64* __restrict__ boards, int64_t n, uint4* __restrict__ out, int64_t rows, int64_t n_cols) {
    constexpr int V = OneHotChunk<T>::V;
    constexpr int CPB = 496 / V;  // chunks per board
    const int64_t total = n * CPB;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t bi = g / CPB;
        const int q = (int)(g - bi * CPB);
        int64_t src = bi;
        if (rows > 0) {  // out is env-major (col, row); records are time-major (row, col)
            const int64_t col = bi / rows;
            const int64_t row = bi - col * rows;
            src = row * n_cols + col;
        }
        const u64 b = __ldg(&boards[src]);
        const int j0 = q * V;
        const int c0 = j0 / 31;
        const int e0 = (int)(b >> (4 * c0)) & 15;
        const int c1 = min(c0 + 1, 15);
        const int e1 = (int)(b >> (4 * c1)) & 15;
        int rel0 = 31 * c0 + e0 - j0;
        int rel1 = (c0 < 15) ? (31 * (c0 + 1) + e1 - j0) : -1;
        if (rel0 < 0 || rel0 >= V) rel0 = -64;
        if (rel1 < 0 || rel1 >= V) rel1 = -64;
        __stcs(&out[g], OneHotChunk<T>::make(rel0, rel1));
    }
}
#endif  // G2048_LEGACY_KERNELS

// ------------------------------------------------------------------------------------------------
// (T,B) time-major records -> (B,T) env-major reference arrays, 32x32 tiles through shared memory
// so that both the reads (along B) and the writes (along T) are coalesced.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
unpack_records_kernel(const uint8_t* __restrict__ meta, const float* __restrict__ rewards,
                      const float* __restrict__ log_probs, const float* __restrict__ values, int64_t t_steps, int64_t n,
                      int32_t* __restrict__ o_actions, uchar4* __restrict__ o_masks, uint8_t* __restrict__ o_term,
                      float* __restrict__ o_rewards, float* __restrict__ o_log_probs, float* __restrict__ o_values) {
    __shared__ uint8_t s_meta[32][33];
    __shared__ float s_r[32][33], s_lp[32][33], s_v[32][33];
    const int64_t e0 = (int64_t)blockIdx.x * 32, t0 = (int64_t)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int k = ty; k < 32; k += 8) {
        const int64_t t = t0 + k, e = e0 + tx;
        if (t < t_steps && e < n) {
            const int64_t i = t * n + e;
            s_meta[k][tx] = meta ? meta[i] : 0;
            if (rewards) s_r[k][tx] = rewards[i];
            if (log_probs) s_lp[k][tx] = log_probs[i];
            if (values) s_v[k][tx] = values[i];
        }
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int64_t e = e0 + k, t = t0 + tx;
        if (t < t_steps && e < n) {
            const int64_t o = e * t_steps + t;
            const uint32_t m = s_meta[tx][k];
            if (o_actions) o_actions[o] = (int32_t)(m & 3u);
            if (o_masks) o_masks[o] = make_uchar4((m >> 2) & 1u, (m >> 3) & 1u, (m >> 4) & 1u, (m >> 5) & 1u);
            if (o_term) o_term[o] = (uint8_t)((m >> 6) & 1u);
            if (o_rewards && rewards) o_rewards[o] = s_r[tx][k];
            if (o_log_probs && log_probs) o_log_probs[o] = s_lp[tx][k];
            if (o_values && values) o_values[o] = s_v[tx][k];
        }
    }
}

// first done + 1 per env, 0 if the env never terminated within t_steps (rollout_buffer.py:168-175)
__global__ void __launch_bounds__(256)
episode_lengths_kernel(const uint8_t* __restrict__ meta, int64_t t_steps, int64_t n, uint32_t* __restrict__ lengths) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    // 16 steps per round trip: with a test after every load the walk is one dependent memory latency per step
    // (150 us for 316 steps x 4 096 envs -- the envs are all the parallelism a small batch has)
    uint32_t len = 0;
    for (int64_t t0 = 0; t0 < t_steps; t0 += 16) {
        uint32_t done_bits = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int64_t t = t0 + k;
            const uint32_t m = (t < t_steps) ? (uint32_t)__ldg(&meta[t * n + e]) : 0u;
            done_bits |= ((m >> 6) & 1u) << k;
        }
        if (done_bits) {
            len = (uint32_t)t0 + (uint32_t)__ffs((int)done_bits);
            break;
        }
    }
    lengths[e] = len;
}

// Exclusive scan of n uint32 (episode lengths) into n + 1 int64 offsets, three small launches and no scratch memory:
// chunk sums are parked where the output needs them anyway -- out[(c + 1) * CHUNK] is by definition the sum of
// chunks 0..c -- one CTA turns them into those running sums, and every chunk then scans its own elements from its
// base out[c * CHUNK].  (The first form was ONE CTA walking the array 1 024 elements per iteration, four barriers and
// an unprefetched load each: ~1.5 us per iteration, 0.4 ms for C4's 262 144 envs -- more than the compaction it feeds.)
constexpr int SCAN_CHUNK = 2048;  // elements per CTA of 256 threads

__device__ __forceinline__ long long warp_inclusive_scan(long long x, int lane) {
    for (int off = 1; off < 32; off <<= 1) {
        const long long y = __shfl_up_sync(0xFFFFFFFFu, x, off);
        if (lane >= off) x += y;
    }
    return x;
}

__global__ void __launch_bounds__(256)
scan_chunk_sums_kernel(const uint32_t* __restrict__ in, int64_t n, long long* __restrict__ out) {
    __shared__ long long s_warp[8];
    const int64_t lo = (int64_t)blockIdx.x * SCAN_CHUNK;
    const int64_t hi = lo + SCAN_CHUNK < n ? lo + SCAN_CHUNK : n;
    long long sum = 0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += 256) sum += (long long)in[i];
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_down_sync(0xFFFFFFFFu, sum, off);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long total = 0;
        for (int w = 0; w < 8; ++w) total += s_warp[w];
        out[hi] = total;  // hi = (c + 1) * CHUNK, or n for the last chunk
    }
}

// one CTA: out[pos(j)] (the chunk sums) -> inclusive running sums, pos(j) = min((j + 1) * CHUNK, n); out[0] = 0
__global__ void __launch_bounds__(1024)
scan_chunk_bases_kernel(long long* __restrict__ out, int64_t n, int64_t chunks) {
    __shared__ long long s_warp[32];
    __shared__ long long s_carry;
    if (threadIdx.x == 0) {
        s_carry = 0;
        out[0] = 0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < chunks; base += 1024) {
        const int64_t j = base + threadIdx.x;
        const int64_t pos = (j + 1) * SCAN_CHUNK < n ? (j + 1) * SCAN_CHUNK : n;
        const long long x = (j < chunks) ? out[pos] : 0;
        const long long incl = warp_inclusive_scan(x, lane);
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) s_warp[lane] = warp_inclusive_scan(s_warp[lane], lane);
        __syncthreads();
        const long long mine = s_carry + (warp ? s_warp[warp - 1] : 0) + incl;
        if (j < chunks) out[pos] = mine;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = mine;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
scan_chunks_kernel(const uint32_t* __restrict__ in, int64_t n, long long* __restrict__ out) {
    __shared__ long long s_warp[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t lo = (int64_t)blockIdx.x * SCAN_CHUNK;
    const long long base = out[lo];  // written by scan_chunk_bases_kernel; this CTA rewrites it with the same value
    // warp w takes 256 consecutive elements, lane l those at 32 k + l (coalesced): eight warp scans with a running total
    const int64_t w_lo = lo + warp * 256;
    long long before[8], running = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int64_t i = w_lo + 32 * k + lane;
        const long long x = (i < n) ? (long long)in[i] : 0;
        const long long incl = warp_inclusive_scan(x, lane);
        before[k] = running + incl - x;
        running += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    if (lane == 0) s_warp[warp] = running;
    __syncthreads();
    long long prefix = base;
    for (int w = 0; w < warp; ++w) prefix += s_warp[w];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int64_t i = w_lo + 32 * k + lane;
        if (i < n) out[i] = prefix + before[k];
    }
}

// ragged compaction: flat[out_base + offsets[e] + t] = rec[t][e] for t < lengths[e].  The records are time-major
// (consecutive envs are adjacent), the buffer env-major (consecutive steps are adjacent): a transpose.  A CTA moves a
// tile of 32 envs x 32 steps through shared memory -- rows read along the envs, columns written along the steps --
// so both sides are whole segments; dead steps (t >= length) are neither read nor written and a tile without a live
// step returns at once.  (The first form gave each thread one env and let it write its steps one by one: stores
// 118 elements apart across a warp, 0.7 TB/s on C4's 262 144 x 390 records.)
__global__ void __launch_bounds__(256)
compact_records_kernel(const u64* __restrict__ rec_boards, const uint8_t* __restrict__ rec_meta,
                       const float* __restrict__ rec_rewards, const float* __restrict__ rec_log_probs,
                       const float* __restrict__ rec_values, int64_t t_steps, int64_t n,
                       const uint32_t* __restrict__ lengths, const long long* __restrict__ offsets, int64_t out_base,
                       u64* __restrict__ boards, uint8_t* __restrict__ meta, float* __restrict__ rewards,
                       float* __restrict__ log_probs, float* __restrict__ values) {
    __shared__ u64 s_boards[32][33];
    __shared__ float s_rewards[32][33], s_log_probs[32][33], s_values[32][33];
    __shared__ uint8_t s_meta[32][36];
    __shared__ uint32_t s_len[32];
    __shared__ long long s_dst[32];
    __shared__ uint32_t s_any;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t e0 = (int64_t)blockIdx.x * 32, t0 = (int64_t)blockIdx.y * 32;
    if (warp == 0) {
        const int64_t e = e0 + lane;
        const uint32_t len = e < n ? (uint32_t)min((int64_t)lengths[e], t_steps) : 0u;  // never past the recorded steps
        s_len[lane] = len;
        s_dst[lane] = e < n ? out_base + offsets[e] : 0;
        const unsigned live = __ballot_sync(0xFFFFFFFFu, (int64_t)len > t0);
        if (lane == 0) s_any = live;
    }
    __syncthreads();
    if (s_any == 0u) return;  // every env of the tile ended before its first step
    const bool has_lp = log_probs && rec_log_probs, has_v = values && rec_values;
    {   // rows of the tile: lane = env
        const int64_t e = e0 + lane;
        const int64_t len = s_len[lane];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = warp + 8 * k;
            const int64_t t = t0 + r;
            if (t < len) {  // implies e < n and t < t_steps
                const int64_t i = t * n + e;
                if (boards) s_boards[r][lane] = rec_boards[i];
                if (meta) s_meta[r][lane] = rec_meta[i];
                if (rewards) s_rewards[r][lane] = rec_rewards[i];
                if (has_lp) s_log_probs[r][lane] = rec_log_probs[i];
                if (has_v) s_values[r][lane] = rec_values[i];
            }
        }
    }
    __syncthreads();
    // columns of the tile: lane = step
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = warp + 8 * k;
        const int64_t t = t0 + lane;
        if (t < (int64_t)s_len[c]) {
            const int64_t o = s_dst[c] + t;
            if (boards) boards[o] = s_boards[lane][c];
            if (meta) meta[o] = s_meta[lane][c];
            if (rewards) rewards[o] = s_rewards[lane][c];
            if (has_lp) log_probs[o] = s_log_probs[lane][c];
            if (has_v) values[o] = s_values[lane][c];
        }
    }
}

__global__ void __launch_bounds__(256)
unpack_flat_meta_kernel(const uint8_t* __restrict__ meta, int64_t n, float4* __restrict__ actions_onehot,
                        uchar4* __restrict__ masks, uint8_t* __restrict__ term) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t m = meta[i];
    const uint32_t a = m & 3u;
    if (actions_onehot)
        actions_onehot[i] = make_float4(a == 0 ? 1.f : 0.f, a == 1 ? 1.f : 0.f, a == 2 ? 1.f : 0.f, a == 3 ? 1.f : 0.f);
    if (masks) masks[i] = make_uchar4((m >> 2) & 1u, (m >> 3) & 1u, (m >> 4) & 1u, (m >> 5) & 1u);
    if (term) term[i] = (uint8_t)((m >> 6) & 1u);
}

// status byte -> pgx State fields: legal_action_mask (n,4) and terminated (n)
__global__ void __launch_bounds__(256)
unpack_status_kernel(const uint8_t* __restrict__ status, int64_t n, uchar4* __restrict__ masks,
                     uint8_t* __restrict__ term) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t m = status[i];
    if (masks) masks[i] = make_uchar4(m & 1u, (m >> 1) & 1u, (m >> 2) & 1u, (m >> 3) & 1u);
    if (term) term[i] = (uint8_t)((m >> 4) & 1u);
}

// one-hot observation (n,16,31) -> bitboard: the inverse of expand_obs, i.e. the
// `observations.argmax(-1)` of src/runs/run_actions_max_tile.py:61-63.  One thread per cell,
// 16 consecutive lanes assemble one board with shuffles.
template <typename T>
__global__ void __launch_bounds__(256)
pack_obs_kernel(const T* __restrict__ obs, int64_t n, u64* __restrict__ boards) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // cell index
    const int64_t total = n * 16;
    u64 part = 0;
    if (g < total) {
        const T* cell = obs + g * 31;
        int best = 0;
        float best_v = (float)cell[0];
        for (int c = 1; c < 31; ++c) {
            const float v = (float)cell[c];
            if (v > best_v) {  // argmax, first index on ties
                best_v = v;
                best = c;
            }
        }
        part = (u64)(best & 15) << (4 * (int)(g & 15));
    }
    for (int off = 8; off > 0; off >>= 1) part |= __shfl_xor_sync(0xFFFFFFFFu, part, off);
    if (g < total && (g & 15) == 0) boards[g >> 4] = part;
}

// per feature row: (count, mean, population variance) in fp64, two passes like numpy's mean / var
__global__ void __launch_bounds__(1024)
row_moments_kernel(const double* __restrict__ x, int64_t n, double* __restrict__ out) {
    __shared__ double s_red[32];
    __shared__ double s_mean;
    const double* row = x + (int64_t)blockIdx.x * n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    auto block_sum = [&](double v) -> double {
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, off);
        __syncthreads();
        if (lane == 0) s_red[warp] = v;
        __syncthreads();
        double t = 0.0;
        if (warp == 0) {
            t = s_red[lane];
            for (int off = 16; off > 0; off >>= 1) t += __shfl_down_sync(0xFFFFFFFFu, t, off);
        }
        return t;  // valid in thread 0
    };
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += row[i];
    const double total = block_sum(acc);
    if (threadIdx.x == 0) s_mean = total / (double)n;
    __syncthreads();
    const double mean = s_mean;
    acc = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const double d = row[i] - mean;
        acc += d * d;
    }
    const double ss = block_sum(acc);
    if (threadIdx.x == 0) {
        out[3 * blockIdx.x + 0] = (double)n;
        out[3 * blockIdx.x + 1] = mean;
        out[3 * blockIdx.x + 2] = ss / (double)n;
    }
}


// ------------------------------------------------------------------------------------------------
// Keyed pseudo-random bijection of [0, n) evaluated at m points: the random subset of the buffer an
// epoch trains on and the epoch's shuffle (torch.randperm(total_length)[:length], src/ppo/data_loader.py:73-101)
// in O(m) instead of a sort of all n positions.  Balanced Feistel network, 4 rounds, over 2h bits with
// 4^h >= n (so 4^h < 4n); round function = low h bits of the first word of Threefry-2x32(key; (right half,
// round)); images >= n go through the network again (cycle walking, < 4 passes expected), which restricts the
// bijection of [0, 4^h) to one of [0, n).  Integer work, 8 B written per index.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
random_subset_kernel(Key k, uint64_t n, int h, int64_t first, int64_t m, int64_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    out[i] = (int64_t)feistel_position(k, (uint64_t)(first + i), h, n);
}

// ------------------------------------------------------------------------------------------------
// RolloutBuffer.store_batch for ARBITRARY per-step rows (rollout_buffer.py:128-187 is generic in the observation
// and action shapes; only 2048 one-hot observations can be packed as bitboards).  Sources are env-major
// (n_envs, t_steps, row_bytes): the kept steps 0..first_done of an env are one contiguous run in the source and in
// the flat destination, so the compaction is one segment copy per env -- row_bytes * (1 read + 1 write) per kept step.
// ------------------------------------------------------------------------------------------------
// first nonzero flag + 1 per env, 0 if there is none; one warp per env, 32 steps per ballot
__global__ void __launch_bounds__(256)
first_done_rows_kernel(const uint8_t* __restrict__ term, int64_t n_envs, int64_t t_steps, uint32_t* __restrict__ lengths) {
    const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (e >= n_envs) return;  // warp-uniform
    const uint8_t* row = term + e * t_steps;
    uint32_t len = 0;
    for (int64_t base = 0; base < t_steps; base += 32) {
        const int64_t t = base + lane;
        const unsigned hit = __ballot_sync(0xFFFFFFFFu, t < t_steps && row[t] != 0);
        if (hit) {
            len = (uint32_t)(base + __ffs((int)hit));
            break;
        }
    }
    if (lane == 0) lengths[e] = len;
}

template <typename V>
__global__ void __launch_bounds__(256)
compact_rows_kernel(const V* __restrict__ src, int64_t env_stride, int64_t row_units, const uint32_t* __restrict__ lengths,
                    const long long* __restrict__ offsets, int64_t out_base, V* __restrict__ dst) {
    const int64_t e = blockIdx.x;
    const int64_t units = (int64_t)lengths[e] * row_units;  // V-sized units to copy for this env
    const V* s = src + e * env_stride;
    V* d = dst + (out_base + offsets[e]) * row_units;
    for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < units; i += (int64_t)gridDim.y * blockDim.x)
        d[i] = __ldcs(&s[i]);
}
}  // namespace g2048

using namespace g2048;

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

#ifdef G2048_LEGACY_KERNELS
extern "C" int g2048_expand_obs_v1(const uint64_t* d_boards, int64_t n, int dtype, void* d_out, int64_t rows,
                                int64_t n_cols, void* stream) {
    G2048_REQUIRE(n >= 0 && rows >= 0 && (rows == 0 || (n_cols > 0 && rows * n_cols == n)), "expand_obs: shape");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_boards && d_out && aligned16(d_out), "expand_obs: pointers (out must be 16-byte aligned)");
    const int sms = sm_count();
    if (sms <= 0) return fail_arg("expand_obs: no device");
    cudaStream_t st = (cudaStream_t)stream;
    auto grid_for = [&](int64_t chunks) {
        int64_t g = (chunks + 255) / 256;
        const int64_t cap = (int64_t)sms * 8 * 4;  // a few waves of 8 resident CTAs per SM
        return (unsigned)(g < cap ? g : cap);
    };
    switch (dtype) {
        case G2048_OBS_F32:
            expand_obs_kernel<float><<<grid_for(n * 124), 256, 0, st>>>((const u64*)d_boards, n, (uint4*)d_out, rows, n_cols);
            break;
        case G2048_OBS_BF16:
            expand_obs_kernel<__nv_bfloat16><<<grid_for(n * 62), 256, 0, st>>>((const u64*)d_boards, n, (uint4*)d_out, rows, n_cols);
            break;
        case G2048_OBS_BOOL:
            expand_obs_kernel<uint8_t><<<grid_for(n * 31), 256, 0, st>>>((const u64*)d_boards, n, (uint4*)d_out, rows, n_cols);
            break;
        default:
            return fail_arg("expand_obs: dtype");
    }
    G2048_CHECK_LAUNCH("expand_obs");
    return G2048_OK;
}
#endif  // G2048_LEGACY_KERNELS

extern "C" int g2048_unpack_records(const uint8_t* d_rec_meta, const float* d_rec_rewards,
                                    const float* d_rec_log_probs, const float* d_rec_values, int64_t t_steps, int64_t n,
                                    int32_t* d_actions, uint8_t* d_masks, uint8_t* d_terminations, float* d_rewards,
                                    float* d_log_probs, float* d_values, void* stream) {
    G2048_REQUIRE(t_steps >= 0 && n >= 0, "unpack_records: shape");
    if (t_steps == 0 || n == 0) return G2048_OK;
    G2048_REQUIRE(d_rec_meta || !(d_actions || d_masks || d_terminations), "unpack_records: meta");
    const dim3 grid((unsigned)((n + 31) / 32), (unsigned)((t_steps + 31) / 32));
    G2048_REQUIRE(grid.y <= 65535u, "unpack_records: too many steps");
    unpack_records_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_rec_meta, d_rec_rewards, d_rec_log_probs,
                                                                 d_rec_values, t_steps, n, d_actions, (uchar4*)d_masks,
                                                                 d_terminations, d_rewards, d_log_probs, d_values);
    G2048_CHECK_LAUNCH("unpack_records");
    return G2048_OK;
}

extern "C" int g2048_episode_lengths(const uint8_t* d_rec_meta, int64_t t_steps, int64_t n, uint32_t* d_lengths,
                                     void* stream) {
    G2048_REQUIRE(t_steps >= 0 && n >= 0, "episode_lengths: shape");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_lengths && (d_rec_meta || t_steps == 0), "episode_lengths: pointers");
    episode_lengths_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(d_rec_meta, t_steps, n, d_lengths);
    G2048_CHECK_LAUNCH("episode_lengths");
    return G2048_OK;
}

extern "C" int g2048_exclusive_scan(const uint32_t* d_in, int64_t n, int64_t* d_out, void* stream) {
    G2048_REQUIRE(n >= 0 && d_out && (d_in || n == 0), "exclusive_scan");
    const int64_t chunks = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
    G2048_REQUIRE(chunks <= 0x7FFFFFFFll, "exclusive_scan: too long");
    cudaStream_t st = (cudaStream_t)stream;
    if (chunks) scan_chunk_sums_kernel<<<(unsigned)chunks, 256, 0, st>>>(d_in, n, (long long*)d_out);
    scan_chunk_bases_kernel<<<1, 1024, 0, st>>>((long long*)d_out, n, chunks);
    if (chunks) scan_chunks_kernel<<<(unsigned)chunks, 256, 0, st>>>(d_in, n, (long long*)d_out);
    G2048_CHECK_LAUNCH("exclusive_scan");
    return G2048_OK;
}

extern "C" int g2048_compact_records(const uint64_t* d_rec_boards, const uint8_t* d_rec_meta,
                                     const float* d_rec_rewards, const float* d_rec_log_probs,
                                     const float* d_rec_values, int64_t t_steps, int64_t n, const uint32_t* d_lengths,
                                     const int64_t* d_offsets, int64_t out_base, uint64_t* d_boards, uint8_t* d_meta,
                                     float* d_rewards, float* d_log_probs, float* d_values, void* stream) {
    G2048_REQUIRE(t_steps >= 0 && n >= 0 && out_base >= 0, "compact_records: shape");
    if (t_steps == 0 || n == 0) return G2048_OK;
    G2048_REQUIRE(d_lengths && d_offsets, "compact_records: lengths/offsets");
    G2048_REQUIRE((!d_boards || d_rec_boards) && (!d_meta || d_rec_meta) && (!d_rewards || d_rec_rewards),
                  "compact_records: sources");
    const dim3 grid(blocks_for(n, 32), (unsigned)((t_steps + 31) / 32));
    G2048_REQUIRE(grid.y <= 65535u, "compact_records: too many steps");
    compact_records_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const u64*)d_rec_boards, d_rec_meta, d_rec_rewards, d_rec_log_probs, d_rec_values, t_steps, n, d_lengths,
        (const long long*)d_offsets, out_base, (u64*)d_boards, d_meta, d_rewards, d_log_probs, d_values);
    G2048_CHECK_LAUNCH("compact_records");
    return G2048_OK;
}

extern "C" int g2048_unpack_flat_meta(const uint8_t* d_meta, int64_t n, float* d_actions_onehot, uint8_t* d_masks,
                                      uint8_t* d_terminations, void* stream) {
    G2048_REQUIRE(n >= 0, "unpack_flat_meta: n");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_meta && (!d_actions_onehot || aligned16(d_actions_onehot)), "unpack_flat_meta: pointers");
    unpack_flat_meta_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
        d_meta, n, (float4*)d_actions_onehot, (uchar4*)d_masks, d_terminations);
    G2048_CHECK_LAUNCH("unpack_flat_meta");
    return G2048_OK;
}

extern "C" int g2048_unpack_status(const uint8_t* d_status, int64_t n, uint8_t* d_masks, uint8_t* d_terminated,
                                   void* stream) {
    G2048_REQUIRE(n >= 0, "unpack_status: n");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_status, "unpack_status: pointers");
    unpack_status_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(d_status, n, (uchar4*)d_masks,
                                                                              d_terminated);
    G2048_CHECK_LAUNCH("unpack_status");
    return G2048_OK;
}

extern "C" int g2048_pack_obs(const void* d_obs, int dtype, int64_t n, uint64_t* d_boards, void* stream) {
    G2048_REQUIRE(n >= 0, "pack_obs: n");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_obs && d_boards, "pack_obs: pointers");
    const unsigned g = blocks_for(n * 16, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == G2048_OBS_BOOL) pack_obs_kernel<uint8_t><<<g, 256, 0, st>>>((const uint8_t*)d_obs, n, (u64*)d_boards);
    else if (dtype == G2048_OBS_F32) pack_obs_kernel<float><<<g, 256, 0, st>>>((const float*)d_obs, n, (u64*)d_boards);
    else return fail_arg("pack_obs: dtype");
    G2048_CHECK_LAUNCH("pack_obs");
    return G2048_OK;
}

extern "C" int g2048_row_moments(const double* d_x, int64_t n_features, int64_t n, double* d_out, void* stream) {
    G2048_REQUIRE(n_features >= 0 && n > 0 && n_features <= 0x7FFFFFFF, "row_moments: shape");
    if (n_features == 0) return G2048_OK;
    G2048_REQUIRE(d_x && d_out, "row_moments: pointers");
    row_moments_kernel<<<(unsigned)n_features, 1024, 0, (cudaStream_t)stream>>>(d_x, n, d_out);
    G2048_CHECK_LAUNCH("row_moments");
    return G2048_OK;
}

extern "C" int g2048_random_subset(uint32_t key0, uint32_t key1, int64_t n, int64_t first, int64_t m, int64_t* d_out,
                                   void* stream) {
    G2048_REQUIRE(n >= 0 && first >= 0 && m >= 0 && first + m <= n && n <= (1ll << 62), "random_subset: range");
    if (m == 0) return G2048_OK;
    G2048_REQUIRE(d_out, "random_subset: pointers");
    const int h = feistel_half_bits((uint64_t)n);  // 4^h >= n
    random_subset_kernel<<<blocks_for(m, 256), 256, 0, (cudaStream_t)stream>>>(Key{key0, key1}, (uint64_t)n, h, first, m,
                                                                              d_out);
    G2048_CHECK_LAUNCH("random_subset");
    return G2048_OK;
}

extern "C" int g2048_first_done_rows(const uint8_t* d_terminations, int64_t n_envs, int64_t t_steps, uint32_t* d_lengths,
                                     void* stream) {
    G2048_REQUIRE(n_envs >= 0 && t_steps >= 0, "first_done_rows: shape");
    if (n_envs == 0) return G2048_OK;
    G2048_REQUIRE(d_lengths && (d_terminations || t_steps == 0), "first_done_rows: pointers");
    first_done_rows_kernel<<<blocks_for(n_envs * 32, 256), 256, 0, (cudaStream_t)stream>>>(d_terminations, n_envs, t_steps,
                                                                                          d_lengths);
    G2048_CHECK_LAUNCH("first_done_rows");
    return G2048_OK;
}

extern "C" int g2048_compact_rows(const void* d_src, int64_t n_envs, int64_t t_steps, int64_t row_bytes,
                                  const uint32_t* d_lengths, const int64_t* d_offsets, int64_t out_base, void* d_dst,
                                  void* stream) {
    G2048_REQUIRE(n_envs >= 0 && t_steps >= 0 && row_bytes > 0 && out_base >= 0 && n_envs <= 0x7FFFFFFFll, "compact_rows: shape");
    if (n_envs == 0 || t_steps == 0) return G2048_OK;
    G2048_REQUIRE(d_src && d_dst && d_lengths && d_offsets, "compact_rows: pointers");
    // widest unit that divides the row and both base addresses (segments start at multiples of the row size)
    const uintptr_t bits = (uintptr_t)d_src | (uintptr_t)d_dst | (uintptr_t)row_bytes;
    const int64_t env_bytes = t_steps * row_bytes;
    const unsigned chunks = (unsigned)std::min<int64_t>(64, (env_bytes / 16 + 255) / 256 + 1);  // CTAs per env
    const dim3 grid((unsigned)n_envs, chunks);
    cudaStream_t st = (cudaStream_t)stream;
    if ((bits & 15u) == 0)
        compact_rows_kernel<uint4><<<grid, 256, 0, st>>>((const uint4*)d_src, env_bytes / 16, row_bytes / 16, d_lengths,
                                                         (const long long*)d_offsets, out_base, (uint4*)d_dst);
    else if ((bits & 3u) == 0)
        compact_rows_kernel<uint32_t><<<grid, 256, 0, st>>>((const uint32_t*)d_src, env_bytes / 4, row_bytes / 4, d_lengths,
                                                            (const long long*)d_offsets, out_base, (uint32_t*)d_dst);
    else
        compact_rows_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t*)d_src, env_bytes, row_bytes, d_lengths,
                                                           (const long long*)d_offsets, out_base, (uint8_t*)d_dst);
    G2048_CHECK_LAUNCH("compact_rows");
    return G2048_OK;
}
