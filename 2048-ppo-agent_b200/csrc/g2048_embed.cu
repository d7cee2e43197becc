// Packed boards straight into the policy's input embedding.
//
// The reference feeds one-hot observations (B,16,31) through Linear(31 -> d_model, bias=False)
// (src/ppo/ppo_agent.py:60,108).  A one-hot row times W^T is row `exponent` of W^T, so the 1 984-byte float
// observation never has to exist: the forward is a row gather from a table that lives in shared memory, the
// backward a segmented sum of the output gradient into the (at most 16) rows a nibble can name.
//
//   forward  : the 31 x d_model table is staged in shared memory once per CTA; every cell's output row is one
//              shared -> global bulk copy (cp.async.bulk, SASS UBLKCP) issued straight from the table.  The
//              threads execute ~12 instructions per row; HBM-write bound (d_model x itemsize bytes per cell).
//   backward : column-owning threads accumulate rows into per-group fp32 tables in shared memory (no atomics),
//              CTAs write partial tables, a second kernel adds them in a fixed order: deterministic.
#include <cuda_bf16.h>

#include "g2048_common.cuh"
#include "g2048_tma.cuh"

namespace g2048 {

typedef unsigned long long u64;

constexpr int EMBED_ROWS = 31;        // OBS_DIM (src/env_definitions.py:2)
constexpr int EMBED_LIVE_ROWS = 16;   // a nibble names exponents 0..15 only
constexpr int EMBED_THREADS = 256;

__global__ void __launch_bounds__(EMBED_THREADS)
embed_boards_kernel(const u64* __restrict__ boards, int64_t n, const uint4* __restrict__ table, int row_bytes,
                    uint8_t* __restrict__ out, const int64_t* __restrict__ indices) {
    extern __shared__ __align__(128) uint8_t s_table[];
    const int chunks = EMBED_ROWS * row_bytes / 16;
    for (int i = threadIdx.x; i < chunks; i += EMBED_THREADS) reinterpret_cast<uint4*>(s_table)[i] = __ldg(&table[i]);
    fence_proxy_async();  // generic-proxy writes of the table -> visible to the bulk copies below
    __syncthreads();

    const int64_t cells = n * 16;
    const int64_t stride = (int64_t)gridDim.x * EMBED_THREADS;
    for (int64_t c = (int64_t)blockIdx.x * EMBED_THREADS + threadIdx.x; c < cells; c += stride) {
        int64_t b = c >> 4;
        if (indices) b = __ldg(&indices[b]);
        const int e = (int)((__ldg(&boards[b]) >> (4 * (int)(c & 15))) & 15ull);
        bulk_store(out + c * row_bytes, s_table + e * row_bytes, (uint32_t)row_bytes);
    }
    bulk_commit();
    bulk_wait_read<0>();  // the table must outlive every copy that reads it
}

// Plain-store forward: a warp copies a row with 16-byte shared loads and streaming global stores.
__global__ void __launch_bounds__(EMBED_THREADS)
embed_boards_plain_kernel(const u64* __restrict__ boards, int64_t n, const uint4* __restrict__ table, int row_bytes,
                       uint8_t* __restrict__ out, const int64_t* __restrict__ indices) {
    extern __shared__ __align__(128) uint8_t s_table[];
    const int chunks = EMBED_ROWS * row_bytes / 16;
    for (int i = threadIdx.x; i < chunks; i += EMBED_THREADS) reinterpret_cast<uint4*>(s_table)[i] = __ldg(&table[i]);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int row_chunks = row_bytes / 16;
    const int64_t cells = n * 16;
    const int64_t warps = (int64_t)gridDim.x * (EMBED_THREADS / 32);
    for (int64_t c = (int64_t)blockIdx.x * (EMBED_THREADS / 32) + (threadIdx.x >> 5); c < cells; c += warps) {
        int64_t b = c >> 4;
        if (indices) b = __ldg(&indices[b]);
        const int e = (int)((__ldg(&boards[b]) >> (4 * (int)(c & 15))) & 15ull);
        const uint4* src = reinterpret_cast<const uint4*>(s_table + e * row_bytes);
        uint4* dst = reinterpret_cast<uint4*>(out + c * row_bytes);
        for (int i = lane; i < row_chunks; i += 32) __stcs(&dst[i], src[i]);
    }
}

// 16 bytes of one gradient row: loaded raw, widened to fp32 only when added to the shared-memory table.
// A table row is stored so that a warp's 16-byte shared accesses are contiguous: for bf16 (8 columns per thread)
// columns 8i..8i+3 of thread i live at [4i, 4i+4) and columns 8i+4..8i+7 at [d_model/2 + 4i, ...); phys() maps a
// logical column to its place in the row.
template <typename T> struct GradVec;
template <> struct GradVec<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ int phys(int col, int) { return col; }
    static __device__ __forceinline__ void add(float* row, int, const uint4& x) {
        float4 a = *reinterpret_cast<float4*>(row);
        a.x += __uint_as_float(x.x); a.y += __uint_as_float(x.y); a.z += __uint_as_float(x.z); a.w += __uint_as_float(x.w);
        *reinterpret_cast<float4*>(row) = a;
    }
};
template <> struct GradVec<__nv_bfloat16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ int phys(int col, int d_model) {
        return ((col >> 2) & 1) * (d_model >> 1) + (col >> 3) * 4 + (col & 3);
    }
    static __device__ __forceinline__ void add(float* row, int d_model, const uint4& x) {
        float* hi = row + (d_model >> 1);
        float4 a = *reinterpret_cast<float4*>(row), b = *reinterpret_cast<float4*>(hi);
        a.x += __uint_as_float(x.x << 16); a.y += __uint_as_float(x.x & 0xFFFF0000u);
        a.z += __uint_as_float(x.y << 16); a.w += __uint_as_float(x.y & 0xFFFF0000u);
        b.x += __uint_as_float(x.z << 16); b.y += __uint_as_float(x.z & 0xFFFF0000u);
        b.z += __uint_as_float(x.w << 16); b.w += __uint_as_float(x.w & 0xFFFF0000u);
        *reinterpret_cast<float4*>(row) = a;
        *reinterpret_cast<float4*>(hi) = b;
    }
};

// d_model / VEC threads own one cell's row; blockDim.x / (d_model / VEC) groups work on different cells with a
// private 16 x d_model fp32 table each.  partials: [gridDim.x][16][d_model].  The kernel lives on bytes in
// flight, not on arithmetic: the next eight 16-byte loads of a thread are issued before the read-modify-writes
// of the previous eight (a serial chain through shared memory whenever neighbouring cells share an exponent).
struct EmbedGradBatch {
    static constexpr int U = 8;
    uint4 v[U];
    u64 board[U];   // consumed (shifted to the cell's nibble) only at accumulation time, when it has long arrived
    int64_t first;  // cell index of v[0]; v[u] belongs to cell first + u * stride
};

// Branch-free: out-of-range slots load the last cell again and are skipped at accumulation time, so that all
// 2 * U loads of a batch issue back to back (a predicated block per slot made ptxas consume each board word
// right after its load, exposing U global latencies per batch).
template <typename T>
__device__ __forceinline__ void embed_grad_load(EmbedGradBatch& b, int64_t c, int64_t stride, int64_t cells,
                                                const u64* __restrict__ boards, const T* __restrict__ grad, int d_model,
                                                int col, const int64_t* __restrict__ indices) {
    b.first = c;
    int64_t idx[EmbedGradBatch::U];
#pragma unroll
    for (int u = 0; u < EmbedGradBatch::U; ++u) {
        const int64_t cu = min(c + u * stride, cells - 1);
        idx[u] = cu >> 4;
        b.v[u] = __ldcs(reinterpret_cast<const uint4*>(grad + cu * d_model + col));
    }
    if (indices) {
#pragma unroll
        for (int u = 0; u < EmbedGradBatch::U; ++u) idx[u] = __ldg(&indices[idx[u]]);
    }
#pragma unroll
    for (int u = 0; u < EmbedGradBatch::U; ++u) b.board[u] = __ldg(&boards[idx[u]]);
}

template <typename T>
__global__ void __launch_bounds__(512)
embed_grad_partial_kernel(const u64* __restrict__ boards, int64_t n, const T* __restrict__ grad, int d_model,
                          float* __restrict__ partials, const int64_t* __restrict__ indices) {
    constexpr int VEC = GradVec<T>::N;
    constexpr int U = EmbedGradBatch::U;
    extern __shared__ __align__(16) float s_acc[];  // [groups][16][d_model]
    const int lanes = d_model / VEC;
    const int groups = blockDim.x / lanes;
    const int group = threadIdx.x / lanes, col = (threadIdx.x - group * lanes) * VEC;
    const int table_floats = EMBED_LIVE_ROWS * d_model;
    for (int i = threadIdx.x; i < groups * table_floats; i += blockDim.x) s_acc[i] = 0.0f;
    __syncthreads();

    float* mine = s_acc + group * table_floats + GradVec<T>::phys(col, d_model);
    const int64_t cells = n * 16;
    const int64_t stride = (int64_t)gridDim.x * groups;
    int64_t c = (int64_t)blockIdx.x * groups + group;
    EmbedGradBatch cur, nxt;
    embed_grad_load<T>(cur, c, stride, cells, boards, grad, d_model, col, indices);
    while (c < cells) {
        c += U * stride;
        embed_grad_load<T>(nxt, c, stride, cells, boards, grad, d_model, col, indices);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t cu = cur.first + u * stride;
            if (cu < cells) {
                const int e = (int)((cur.board[u] >> (4 * (int)(cu & 15))) & 15ull);
                GradVec<T>::add(mine + e * d_model, d_model, cur.v[u]);
            }
        }
        cur = nxt;
    }
    __syncthreads();
    float* dst = partials + (size_t)blockIdx.x * table_floats;
    for (int i = threadIdx.x; i < table_floats; i += blockDim.x) {
        const int r = i / d_model, at = r * d_model + GradVec<T>::phys(i - r * d_model, d_model);
        float s = 0.0f;
        for (int g = 0; g < groups; ++g) s += s_acc[g * table_floats + at];
        dst[i] = s;
    }
}

// grad_table[31][d_model]: rows < 16 = sum of the CTAs' partial tables, rows >= 16 = 0.  64 columns x 16 slices per
// CTA: slice s adds partials s, s+16, ... in index order, then the slices are added in slice order -- a fixed tree,
// so the result does not depend on scheduling.
constexpr int REDUCE_COLS = 64, REDUCE_SLICES = 16;

__global__ void __launch_bounds__(REDUCE_COLS * REDUCE_SLICES)
embed_grad_reduce_kernel(const float* __restrict__ partials, int n_partials, int d_model, float* __restrict__ grad_table) {
    __shared__ float s_part[REDUCE_SLICES][REDUCE_COLS];
    const int col = threadIdx.x % REDUCE_COLS, slice = threadIdx.x / REDUCE_COLS;
    const int i = blockIdx.x * REDUCE_COLS + col;
    const int table_floats = EMBED_LIVE_ROWS * d_model;
    float acc = 0.0f;
    if (i < table_floats) {
#pragma unroll 4
        for (int p = slice; p < n_partials; p += REDUCE_SLICES) acc += partials[(size_t)p * table_floats + i];
    }
    s_part[slice][col] = acc;
    __syncthreads();
    if (slice == 0 && i < EMBED_ROWS * d_model) {
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < REDUCE_SLICES; ++k) s += s_part[k][col];
        grad_table[i] = s;
    }
}

struct EmbedGradGeometry {
    int lanes, groups, threads, smem, grid;
};

static bool embed_grad_geometry(int64_t n, int d_model, int itemsize, int sms, EmbedGradGeometry* g) {
    const int vec = 16 / itemsize;
    if (d_model <= 0 || d_model % vec) return false;
    g->lanes = d_model / vec;
    if (g->lanes > 512) return false;
    const int table_bytes = EMBED_LIVE_ROWS * d_model * (int)sizeof(float);
    if (table_bytes > 200 * 1024) return false;
    // groups: up to 256 threads per CTA, tables within 64 KiB so that three CTAs share an SM
    int groups = g->lanes >= 256 ? 1 : 256 / g->lanes;
    const int fit = 64 * 1024 / table_bytes;
    if (groups > fit) groups = fit < 1 ? 1 : fit;
    g->groups = groups;
    g->threads = groups * g->lanes;
    g->smem = groups * table_bytes;
    // CTAs an SM holds: ~128 registers per thread, 227 KiB of shared memory (1 KiB reserved per CTA)
    const int by_regs = 65536 / (g->threads * 128), by_smem = 227 * 1024 / (g->smem + 1024);
    int resident = by_regs < by_smem ? by_regs : by_smem;
    resident = resident < 1 ? 1 : (resident > 4 ? 4 : resident);
    const int64_t want = (n * 16 + (int64_t)groups * 32 - 1) / ((int64_t)groups * 32);  // >= 32 cells per group
    const int64_t cap = (int64_t)sms * resident;
    g->grid = (int)(want < 1 ? 1 : (want < cap ? want : cap));
    return true;
}

}  // namespace g2048

using namespace g2048;

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

static int embed_forward(bool bulk, const uint64_t* d_boards, int64_t n, const int64_t* d_indices, const void* d_table,
                         int d_model, int dtype, void* d_out, void* stream) {
    G2048_REQUIRE(n >= 0 && d_model > 0, "embed_boards: shape");
    G2048_REQUIRE(dtype == G2048_OBS_F32 || dtype == G2048_OBS_BF16, "embed_boards: dtype (f32 or bf16)");
    const int row_bytes = d_model * (dtype == G2048_OBS_F32 ? 4 : 2);
    G2048_REQUIRE(row_bytes % 16 == 0, "embed_boards: d_model * itemsize must be a multiple of 16 bytes");
    const int smem = EMBED_ROWS * row_bytes;
    G2048_REQUIRE(smem <= 200 * 1024, "embed_boards: table does not fit in shared memory");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_boards && d_table && d_out && aligned16(d_table) && aligned16(d_out), "embed_boards: pointers (16-byte aligned)");
    const int sms = sm_count();
    if (sms <= 0) return fail_arg("embed_boards: no device");
    auto kernel = bulk ? embed_boards_kernel : embed_boards_plain_kernel;
    static int configured[64][2];  // dynamic shared memory each kernel is already allowed, per device (0: the default 48 KiB)
    int dev = 0;
    G2048_REQUIRE(cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64, "embed_boards: device");
    if (smem > 48 * 1024 && smem > configured[dev][bulk]) {
        int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "embed_boards: smem attribute");
        if (rc) return rc;
        configured[dev][bulk] = smem;
    }
    const int per_cta = bulk ? EMBED_THREADS : EMBED_THREADS / 32;
    const int64_t need = (n * 16 + per_cta - 1) / per_cta;
    const int resident = smem > 100 * 1024 ? 1 : (smem > 56 * 1024 ? 2 : 4);
    const int64_t cap = (int64_t)sms * resident;
    kernel<<<(unsigned)(need < cap ? need : cap), EMBED_THREADS, smem, (cudaStream_t)stream>>>(
        (const u64*)d_boards, n, (const uint4*)d_table, row_bytes, (uint8_t*)d_out, d_indices);
    G2048_CHECK_LAUNCH("embed_boards");
    return G2048_OK;
}

// Measured on B200 (tools/bench_embed.py, tools/probes/embed_flush_probe.py; 2^18 boards, d_model 256; us):
//                        f32 rows (1 KiB)            bf16 rows (512 B)
//                        L2 warm   after L2 flush    L2 warm   after L2 flush
//   bulk copies            704         694             364         352
//   plain 16-byte stores   635         880             469         652
// The plain stores lose 40 % when L2 is full of another kernel's dirty lines (the normal state inside a training
// step); the bulk copies do not care, so they are the product path and the plain kernel stays for A/B runs.
extern "C" int g2048_embed_boards(const uint64_t* d_boards, int64_t n, const int64_t* d_indices, const void* d_table,
                                  int d_model, int dtype, void* d_out, void* stream) {
    return embed_forward(true, d_boards, n, d_indices, d_table, d_model, dtype, d_out, stream);
}

extern "C" int g2048_embed_boards_bulk(const uint64_t* d_boards, int64_t n, const int64_t* d_indices, const void* d_table,
                                       int d_model, int dtype, void* d_out, void* stream) {
    return embed_forward(true, d_boards, n, d_indices, d_table, d_model, dtype, d_out, stream);
}

extern "C" int g2048_embed_boards_plain(const uint64_t* d_boards, int64_t n, const int64_t* d_indices, const void* d_table,
                                        int d_model, int dtype, void* d_out, void* stream) {
    return embed_forward(false, d_boards, n, d_indices, d_table, d_model, dtype, d_out, stream);
}

extern "C" int64_t g2048_embed_grad_scratch_bytes(int64_t n, int d_model, int dtype) {
    EmbedGradGeometry g;
    const int sms = sm_count();
    if (sms <= 0 || (dtype != G2048_OBS_F32 && dtype != G2048_OBS_BF16)) return -1;
    if (!embed_grad_geometry(n, d_model, dtype == G2048_OBS_F32 ? 4 : 2, sms, &g)) return -1;
    return (int64_t)g.grid * EMBED_LIVE_ROWS * d_model * (int64_t)sizeof(float);
}

extern "C" int g2048_embed_boards_grad(const uint64_t* d_boards, int64_t n, const int64_t* d_indices, const void* d_grad_out,
                                       int d_model, int dtype, float* d_grad_table, void* d_scratch, void* stream) {
    G2048_REQUIRE(n >= 0, "embed_boards_grad: shape");
    G2048_REQUIRE(dtype == G2048_OBS_F32 || dtype == G2048_OBS_BF16, "embed_boards_grad: dtype (f32 or bf16)");
    const int sms = sm_count();
    if (sms <= 0) return fail_arg("embed_boards_grad: no device");
    EmbedGradGeometry g;
    G2048_REQUIRE(embed_grad_geometry(n, d_model, dtype == G2048_OBS_F32 ? 4 : 2, sms, &g),
                  "embed_boards_grad: d_model (16-byte rows, at most 512 column groups, tables within shared memory)");
    G2048_REQUIRE(d_grad_table, "embed_boards_grad: grad_table");
    cudaStream_t st = (cudaStream_t)stream;
    int n_partials = 0;
    if (n > 0) {
        G2048_REQUIRE(d_boards && d_grad_out && d_scratch && aligned16(d_grad_out) && aligned16(d_scratch),
                      "embed_boards_grad: pointers (16-byte aligned)");
        int rc;
        if (dtype == G2048_OBS_F32) {
            rc = check_cuda(cudaFuncSetAttribute(embed_grad_partial_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem), "embed_boards_grad: smem attribute");
            if (rc) return rc;
            embed_grad_partial_kernel<float><<<g.grid, g.threads, g.smem, st>>>((const u64*)d_boards, n, (const float*)d_grad_out, d_model, (float*)d_scratch, d_indices);
        } else {
            rc = check_cuda(cudaFuncSetAttribute(embed_grad_partial_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem), "embed_boards_grad: smem attribute");
            if (rc) return rc;
            embed_grad_partial_kernel<__nv_bfloat16><<<g.grid, g.threads, g.smem, st>>>((const u64*)d_boards, n, (const __nv_bfloat16*)d_grad_out, d_model, (float*)d_scratch, d_indices);
        }
        G2048_CHECK_LAUNCH("embed_boards_grad (partials)");
        n_partials = g.grid;
    }
    embed_grad_reduce_kernel<<<blocks_for((int64_t)EMBED_ROWS * d_model, REDUCE_COLS), REDUCE_COLS * REDUCE_SLICES, 0, st>>>((const float*)d_scratch, n_partials, d_model, d_grad_table);
    G2048_CHECK_LAUNCH("embed_boards_grad (reduce)");
    return G2048_OK;
}
