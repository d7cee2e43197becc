// GAE as a segmented affine REVERSE SCAN (BASELINE.json north_star: "warp shuffles handle the reverse-scan GAE over the
// time dimension"; src/ppo/data_loader.py:103-130).  Opt-in companion of g2048_gae_flat: that kernel walks every episode
// with the reference's exact fp32 operation order (bit-identical, 0.65-0.74 of HBM because a walking lane moves no
// bytes); this one re-associates the recurrence and is a pure stream -- results agree within the 1e-5 relative
// tolerance north_star states for GAE (tests/test_gae_scan_gpu.py), not bit for bit.
//
//   gae_t = delta_t + a_t * gae_{t+1},   a_t = done_t ? 0 : gamma * lambda,
//   delta_t = (r_t + gamma * (done_t ? 0 : V_{t+1})) - V_t
//
// i.e. gae_t = f_t(gae_{t+1}) with the affine map f_t(x) = a_t x + delta_t; maps compose associatively
// ((a1,b1) o (a2,b2) = (a1 a2, a1 b2 + b1)), so the suffix compositions F_t = f_t o f_{t+1} o ... are a scan:
//   * a thread composes its 8 consecutive steps serially (registers, 128-bit loads),
//   * a warp scans the 32 thread aggregates with shuffles (5 steps), the 8 warp aggregates go through shared memory,
//   * a persistent CTA owns a contiguous RANGE of 2 048-step tiles and walks it from its end to its start, so the gae
//     entering a tile is the CTA's own carry -- no CTA ever waits for another one.  (Two earlier forms chained the
//     tiles by decoupled look-back, one tile per CTA or by ticket: every tile's last warps then wait for the
//     neighbour's publication and the rest of the CTA for them -- `barrier` 14.7 stalls per issue, 0.32-0.46 of HBM,
//     profiles/r02_new_kernels.csv.)
//   * what the steps at the END of a range see of the next range is not known while the kernel runs: the carry is kept
//     as an affine map of that unknown x (coefficient 1 at the range's end, exactly 0 after the first `done` or after
//     ~1 700 steps of gamma*lambda = 0.94 underflowing), results are written for x = 0, and a second, tiny kernel --
//     one warp per range -- chains the range aggregates, adds coefficient * x to the few hundred steps whose
//     coefficient is not zero and adds their share of the moments.  No spin-waits anywhere, hence no co-residency
//     requirement.
// 9 B read + 8 B written per step; fp64 moment sums for the normalisation in the same pass.
#include "g2048_common.cuh"
#include "g2048_tma.cuh"

namespace g2048 {

#ifndef G2048_SCAN_THREADS
#define G2048_SCAN_THREADS 256
#endif
#ifndef G2048_SCAN_ITEMS
#define G2048_SCAN_ITEMS 8  // consecutive steps per thread (a multiple of 8)
#endif
#ifndef G2048_SCAN_CTAS
#define G2048_SCAN_CTAS 3  // resident CTAs per SM
#endif
constexpr int SCAN_THREADS = G2048_SCAN_THREADS;
constexpr int SCAN_ITEMS = G2048_SCAN_ITEMS;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 2 048 steps
constexpr int SCAN_WARPS = SCAN_THREADS / 32;
constexpr int SCAN_MAX_RANGES = 4096;

// per range (persistent CTA): the gae of the range's first step as a map of x = the gae of the next range's first step,
// and how many steps at the range's end have a non-zero coefficient of x (written for x = 0, left out of the moments)
struct ScanRange {
    float a, b;
    unsigned int pending;
    unsigned int pad;
};

struct Affine {
    float a, b;
};
// apply `inner` first, then `outer`
__device__ __forceinline__ Affine compose(Affine outer, Affine inner) {
    return Affine{outer.a * inner.a, outer.a * inner.b + outer.b};
}

constexpr int SCAN_STAGES = 3;
// one stage of the shared-memory ring: a tile's rewards, values (+ the 4 values after it) and done flags, as the bulk
// copy engine delivers them
struct __align__(128) ScanStage {
    float r[SCAN_TILE];
    float v[SCAN_TILE + 4];
    uint8_t d[SCAN_TILE];
};
constexpr int SCAN_SMEM_BYTES = SCAN_STAGES * (int)sizeof(ScanStage);

// STAGED: the tiles come in through the bulk copy engine (cp.async.bulk, SASS UBLKCP) into a three-stage ring in shared
// memory, two tiles ahead of the one being scanned, so that a CTA keeps ~36 KB of reads in flight whatever its
// threads are doing -- with loads into registers (also one tile ahead) the in-flight bytes per SM were the limit
// (Little's law: 24 warps x 2.4 KB = 58 KB against a bandwidth-delay product of ~70 KB per SM): 0.67 of HBM.
// Needs 16-byte aligned arrays; the unaligned form (views into a larger buffer) loads through registers.
template <bool STAGED>
__global__ void __launch_bounds__(SCAN_THREADS, G2048_SCAN_CTAS)
gae_scan_kernel(const float* __restrict__ rewards, const float* __restrict__ values, const uint8_t* __restrict__ dones,
                int64_t n, int64_t n_tiles, int64_t tiles_per_range, float gamma, float gamma_lambda, float* __restrict__ adv,
                float* __restrict__ ret, ScanRange* __restrict__ ranges, double* __restrict__ moments) {
    extern __shared__ __align__(128) uint8_t scan_smem[];
    ScanStage* stages = reinterpret_cast<ScanStage*>(scan_smem);
    __shared__ __align__(8) uint64_t s_full[SCAN_STAGES];
    __shared__ Affine s_warp[2][SCAN_WARPS];  // double-buffered by tile parity: one barrier per tile
    __shared__ double s_red[4 * SCAN_WARPS];
    __shared__ unsigned int s_pending;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tile_lo = (int64_t)blockIdx.x * tiles_per_range;
    const int64_t tile_hi = min(n_tiles, tile_lo + tiles_per_range);
    if (threadIdx.x == 0) {
        s_pending = 0u;
        if (STAGED) {
            for (int k = 0; k < SCAN_STAGES; ++k) mbar_init(&s_full[k], 1);
            fence_proxy_async();  // the initialised barriers -> visible to the copy engine
        }
    }
    __syncthreads();

    // a tile that ends inside the buffer with room for the 4 extra values is fetched by three bulk copies; the last
    // tile of the buffer (ragged, or without anything after it) goes through registers
    auto bulk_ok = [&](int64_t tile) { return STAGED && (tile + 1) * SCAN_TILE + 4 <= n; };
    auto issue = [&](int64_t tile, unsigned k) {  // thread 0
        if (tile < tile_lo || !bulk_ok(tile)) return;
        ScanStage& st = stages[k % SCAN_STAGES];
        const int64_t lo = tile * SCAN_TILE;
        mbar_arrive_expect_tx(&s_full[k % SCAN_STAGES], (uint32_t)(SCAN_TILE * 4 + (SCAN_TILE + 4) * 4 + SCAN_TILE));
        bulk_load(st.r, rewards + lo, SCAN_TILE * 4, &s_full[k % SCAN_STAGES]);
        bulk_load(st.v, values + lo, (SCAN_TILE + 4) * 4, &s_full[k % SCAN_STAGES]);
        bulk_load(st.d, dones + lo, SCAN_TILE, &s_full[k % SCAN_STAGES]);
    };
    // the ring serves the tiles from ring_top downwards: use j of the ring (tile ring_top - j) goes through stage j % 3
    // and completes phase j / 3 of that stage's barrier.  Only the last one or two tiles of the BUFFER cannot be
    // fetched in bulk; they sit at the top of their range and are taken through registers before the ring starts.
    int64_t ring_top = tile_hi - 1;
    while (ring_top >= tile_lo && !bulk_ok(ring_top)) --ring_top;
    if (STAGED && threadIdx.x == 0) {
        issue(ring_top, 0);
        issue(ring_top - 1, 1);
    }

    Affine carry{1.0f, 0.0f};  // gae entering the current tile = carry.a * x + carry.b, x = gae of the next range's first step
    double acc4[4] = {0.0, 0.0, 0.0, 0.0};
    unsigned my_pending = 0;
    unsigned it = 0;
    for (int64_t tile = tile_hi - 1; tile >= tile_lo; --tile, ++it) {
        const int64_t first = tile * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;  // this thread's first step
        const bool full = first + SCAN_ITEMS <= n;
        const bool ringed = STAGED && tile <= ring_top;
        const unsigned j = (unsigned)(ring_top - tile);  // use of the ring (meaningful when ringed)
        // two tiles ahead: its stage was last read one iteration ago, before that iteration's barrier
        if (ringed && threadIdx.x == 0) issue(tile - 2, j + 2);

        // ---- 8 consecutive steps per thread ----------------------------------------------------------------------
        float r[SCAN_ITEMS], v[SCAN_ITEMS + 1];
        bool d[SCAN_ITEMS];
        if (ringed) {
            const ScanStage& st = stages[j % SCAN_STAGES];
            mbar_wait(&s_full[j % SCAN_STAGES], (j / SCAN_STAGES) & 1u);
            const int o = threadIdx.x * SCAN_ITEMS;
#pragma unroll
            for (int q = 0; q < SCAN_ITEMS / 4; ++q) {
                const float4 rq = *reinterpret_cast<const float4*>(st.r + o + 4 * q);
                const float4 vq = *reinterpret_cast<const float4*>(st.v + o + 4 * q);
                r[4 * q] = rq.x; r[4 * q + 1] = rq.y; r[4 * q + 2] = rq.z; r[4 * q + 3] = rq.w;
                v[4 * q] = vq.x; v[4 * q + 1] = vq.y; v[4 * q + 2] = vq.z; v[4 * q + 3] = vq.w;
            }
            v[SCAN_ITEMS] = st.v[o + SCAN_ITEMS];
#pragma unroll
            for (int q = 0; q < SCAN_ITEMS / 8; ++q) {
                const uint2 dd = *reinterpret_cast<const uint2*>(st.d + o + 8 * q);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    d[8 * q + k] = ((dd.x >> (8 * k)) & 0xFFu) != 0u;
                    d[8 * q + 4 + k] = ((dd.y >> (8 * k)) & 0xFFu) != 0u;
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) {
                const bool in = first + k < n;
                r[k] = in ? rewards[first + k] : 0.0f;
                v[k] = in ? values[first + k] : 0.0f;
                d[k] = in ? dones[first + k] != 0 : true;  // past the end: a = 0, delta = 0 -- contributes nothing
            }
            v[SCAN_ITEMS] = (first + SCAN_ITEMS < n) ? values[first + SCAN_ITEMS] : 0.0f;  // V of the step after mine (0 past the end)
        }

        // ---- per-step maps and the thread's aggregate (composition of its 8 steps, first step outermost) ---------------
        float a[SCAN_ITEMS], b[SCAN_ITEMS];
        Affine mine{1.0f, 0.0f};
#pragma unroll
        for (int k = SCAN_ITEMS - 1; k >= 0; --k) {
            const float next_v = d[k] ? 0.0f : v[k + 1];
            a[k] = d[k] ? 0.0f : gamma_lambda;
            b[k] = (r[k] + gamma * next_v) - v[k];
            mine = compose(Affine{a[k], b[k]}, mine);
        }
        // ---- warp: inclusive suffix scan over lanes (lane L: composition of lanes L .. 31) ----------------------------
        Affine incl = mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            Affine other;
            other.a = __shfl_down_sync(0xFFFFFFFFu, incl.a, off);
            other.b = __shfl_down_sync(0xFFFFFFFFu, incl.b, off);
            if (lane + off < 32) incl = compose(incl, other);
        }
        if (lane == 0) s_warp[it & 1u][warp] = incl;
        // what follows my steps inside the warp: the inclusive value of the next lane
        Affine after;
        after.a = __shfl_down_sync(0xFFFFFFFFu, incl.a, 1);
        after.b = __shfl_down_sync(0xFFFFFFFFu, incl.b, 1);
        if (lane == 31) after = Affine{1.0f, 0.0f};
        __syncthreads();  // the one barrier of a tile (also: everybody has read this tile's stage)
        // ... the warps after mine, and the whole tile: lane w of every warp takes warp w's aggregate and the (at most 32)
        // aggregates are suffix-scanned with shuffles -- no second barrier, and a fifth of the instructions of every
        // thread composing all of them itself
        Affine agg = (lane < SCAN_WARPS) ? s_warp[it & 1u][lane] : Affine{1.0f, 0.0f};
#pragma unroll
        for (int off = 1; off < SCAN_WARPS; off <<= 1) {
            Affine other;
            other.a = __shfl_down_sync(0xFFFFFFFFu, agg.a, off);
            other.b = __shfl_down_sync(0xFFFFFFFFu, agg.b, off);
            if (lane + off < SCAN_WARPS) agg = compose(agg, other);
        }
        Affine whole, later;  // lane 0 holds warps 0.., lane w + 1 the warps after warp w (identity past the last warp)
        whole.a = __shfl_sync(0xFFFFFFFFu, agg.a, 0);
        whole.b = __shfl_sync(0xFFFFFFFFu, agg.b, 0);
        later.a = __shfl_sync(0xFFFFFFFFu, agg.a, warp + 1);
        later.b = __shfl_sync(0xFFFFFFFFu, agg.b, warp + 1);
        // everything between my last step and the end of the RANGE, as a map of x
        const Affine entering = compose(compose(after, later), carry);
        carry = compose(whole, carry);  // ... and what enters the next (earlier) tile

        // ---- apply: the gae entering my steps for x = 0, then the reference's own recurrence over them -------------------
        float g = entering.b;
        float o_adv[SCAN_ITEMS], o_ret[SCAN_ITEMS];
        float part[4] = {0.0f, 0.0f, 0.0f, 0.0f};  // the thread's 8 steps are summed in fp32, tiles and threads in fp64
        if (entering.a == 0.0f) {  // nothing of x reaches my steps (everywhere but at the end of a range)
#pragma unroll
            for (int k = SCAN_ITEMS - 1; k >= 0; --k) {
                g = b[k] + a[k] * g;
                o_adv[k] = g;
                o_ret[k] = g + v[k];
                if (first + k < n) {
                    part[0] += g;
                    part[1] += g * g;
                    part[2] += o_ret[k];
                    part[3] += o_ret[k] * o_ret[k];
                }
            }
        } else {
            float coef = entering.a;
#pragma unroll
            for (int k = SCAN_ITEMS - 1; k >= 0; --k) {
                g = b[k] + a[k] * g;
                coef = a[k] * coef;  // coefficient of x in this step's gae: non-zero only at the end of the range
                o_adv[k] = g;
                o_ret[k] = g + v[k];
                if (first + k < n) {
                    if (coef != 0.0f) {
                        my_pending += 1u;  // finished by gae_scan_fix_kernel, which also adds it to the moments
                    } else {
                        part[0] += g;
                        part[1] += g * g;
                        part[2] += o_ret[k];
                        part[3] += o_ret[k] * o_ret[k];
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) acc4[k] += (double)part[k];
        if (STAGED && full) {
#pragma unroll
            for (int q = 0; q < SCAN_ITEMS / 4; ++q) {
                *reinterpret_cast<float4*>(adv + first + 4 * q) = make_float4(o_adv[4 * q], o_adv[4 * q + 1], o_adv[4 * q + 2], o_adv[4 * q + 3]);
                *reinterpret_cast<float4*>(ret + first + 4 * q) = make_float4(o_ret[4 * q], o_ret[4 * q + 1], o_ret[4 * q + 2], o_ret[4 * q + 3]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) {
                if (first + k < n) {
                    adv[first + k] = o_adv[k];
                    ret[first + k] = o_ret[k];
                }
            }
        }
    }
    if (my_pending) atomicAdd(&s_pending, my_pending);
    if (moments) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double sum = acc4[k];
            for (int off = 16; off > 0; off >>= 1) sum += __shfl_down_sync(0xFFFFFFFFu, sum, off);
            if (lane == 0) s_red[k * SCAN_WARPS + warp] = sum;
        }
    }
    __syncthreads();
    if (moments && threadIdx.x < 4) {
        double sum = 0.0;
        for (int w = 0; w < SCAN_WARPS; ++w) sum += s_red[threadIdx.x * SCAN_WARPS + w];
        atomicAdd(&moments[1 + threadIdx.x], sum);
    }
    if (threadIdx.x == 0) ranges[blockIdx.x] = ScanRange{carry.a, carry.b, s_pending, 0u};
}

// One CTA per range: x = the gae of the next range's first step (chained through the range aggregates until one does
// not depend on its own successor), then gae += coefficient * x for the `pending` steps at the range's end -- their
// coefficient is (gamma*lambda)^(distance to the range's end + 1): a step with a non-zero coefficient has no `done`
// between itself and the end -- and their share of the moment sums.
constexpr int SCAN_FIX_THREADS = 256;
__global__ void __launch_bounds__(SCAN_FIX_THREADS)
gae_scan_fix_kernel(const ScanRange* __restrict__ ranges, int n_ranges, int64_t n, int64_t tiles_per_range, float gamma_lambda,
                    float* __restrict__ adv, float* __restrict__ ret, double* __restrict__ moments) {
    const int c = blockIdx.x, lane = threadIdx.x & 31;
    const unsigned pending = ranges[c].pending;
    if (threadIdx.x == 0 && c == 0 && moments) atomicAdd(&moments[0], (double)n);
    if (pending == 0u) return;
    Affine acc{1.0f, 0.0f};
    for (int j = c + 1; j < n_ranges && acc.a != 0.0f; ++j) acc = compose(acc, Affine{ranges[j].a, ranges[j].b});
    const float x = acc.b;  // past the last range the gae is 0
    const int64_t end = min(n, (int64_t)(c + 1) * tiles_per_range * SCAN_TILE);  // one past the range's last step
    double acc4[4] = {0.0, 0.0, 0.0, 0.0};
    for (unsigned i = threadIdx.x; i < pending; i += SCAN_FIX_THREADS) {  // a few hundred steps: one or two per thread
        const int64_t t = end - 1 - (int64_t)i;
        const float coef = powf(gamma_lambda, (float)(i + 1u));
        const float g = adv[t] + coef * x, q = ret[t] + coef * x;
        if (x != 0.0f) {
            adv[t] = g;
            ret[t] = q;
        }
        acc4[0] += (double)g;
        acc4[1] += (double)g * (double)g;
        acc4[2] += (double)q;
        acc4[3] += (double)q * (double)q;
    }
    if (moments) {  // (every thread of the CTA gets here: `pending` is uniform)
        __shared__ double s_red[4 * (SCAN_FIX_THREADS / 32)];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double sum = acc4[k];
            for (int off = 16; off > 0; off >>= 1) sum += __shfl_down_sync(0xFFFFFFFFu, sum, off);
            if (lane == 0) s_red[k * (SCAN_FIX_THREADS / 32) + (threadIdx.x >> 5)] = sum;
        }
        __syncthreads();
        if (threadIdx.x < 4) {
            double sum = 0.0;
            for (int w = 0; w < SCAN_FIX_THREADS / 32; ++w) sum += s_red[threadIdx.x * (SCAN_FIX_THREADS / 32) + w];
            atomicAdd(&moments[1 + threadIdx.x], sum);
        }
    }
}

}  // namespace g2048

using namespace g2048;

extern "C" int64_t g2048_gae_scan_scratch_bytes(int64_t n) {
    (void)n;
    return (int64_t)SCAN_MAX_RANGES * (int64_t)sizeof(ScanRange);  // one entry per persistent CTA; need not be zeroed
}

extern "C" int g2048_gae_flat_scan(const float* d_rewards, const float* d_values, const uint8_t* d_dones, int64_t n, double gamma,
                                   double lambda_gae, float* d_adv, float* d_ret, void* d_scan_state, double* d_moments,
                                   void* stream) {
    G2048_REQUIRE(n >= 0, "gae_flat_scan: n");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_rewards && d_values && d_dones && d_adv && d_ret && d_scan_state, "gae_flat_scan: pointers");
    G2048_REQUIRE(((uintptr_t)d_scan_state & 15u) == 0, "gae_flat_scan: scratch must be 16-byte aligned");
    const int64_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    const auto a16 = [](const void* p) { return ((uintptr_t)p & 15u) == 0; };
    const bool aligned = a16(d_rewards) && a16(d_values) && a16(d_adv) && a16(d_ret) && a16(d_dones);  // bulk copies and 128-bit stores
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = sm_count();
    if (sms <= 0) return fail_arg("gae_flat_scan: no device");
    int64_t want = (int64_t)sms * G2048_SCAN_CTAS;  // three 256-thread CTAs per SM (a 54 KiB ring each: two tiles in flight), each walking its own range of tiles
    if (want > SCAN_MAX_RANGES) want = SCAN_MAX_RANGES;
    const int64_t tiles_per_range = (n_tiles + want - 1) / want;
    const int n_ranges = (int)((n_tiles + tiles_per_range - 1) / tiles_per_range);
    const float g = (float)gamma, gl = (float)(gamma * lambda_gae);
    static bool configured_on[64] = {false};
    bool* configured = device_once_flag(configured_on);
    if (!configured) return fail_arg("no CUDA device");
    if (!*configured) {
        const int rc = check_cuda(cudaFuncSetAttribute(gae_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SCAN_SMEM_BYTES),
                                  "gae_flat_scan: shared memory attribute");
        if (rc) return rc;
        *configured = true;
    }
    if (aligned)
        gae_scan_kernel<true><<<n_ranges, SCAN_THREADS, SCAN_SMEM_BYTES, st>>>(d_rewards, d_values, d_dones, n, n_tiles, tiles_per_range,
                                                                              g, gl, d_adv, d_ret, (ScanRange*)d_scan_state, d_moments);
    else
        gae_scan_kernel<false><<<n_ranges, SCAN_THREADS, 0, st>>>(d_rewards, d_values, d_dones, n, n_tiles, tiles_per_range, g, gl,
                                                                  d_adv, d_ret, (ScanRange*)d_scan_state, d_moments);
    G2048_CHECK_LAUNCH("gae_flat_scan");
    gae_scan_fix_kernel<<<n_ranges, SCAN_FIX_THREADS, 0, st>>>((const ScanRange*)d_scan_state, n_ranges, n, tiles_per_range, gl, d_adv, d_ret,
                                                 d_moments);
    G2048_CHECK_LAUNCH("gae_flat_scan: fix-up");
    return G2048_OK;
}
