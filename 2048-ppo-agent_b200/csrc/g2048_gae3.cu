// GAE over a flat buffer, current generation (src/ppo/data_loader.py:103-130; bit-identical results).
//
// What limits this kernel is not arithmetic but how many steps an SM keeps in flight while the serial
// recurrence of the longest episode of a tile runs (one multiply-add per step).  ncu on the earlier
// versions: v1 (1 024-step tiles, everything staged in shared memory) was issue-bound by one-lane
// walks; a 7 680-step version filled by bulk copies kept 9 bytes per step resident, so only 3 tiles
// fitted an SM and HBM idled while they walked (23 % DRAM throughput); a scalar-load version of the
// present layout spent 27 % of its samples waiting on 4-byte loads and 40 % at the barrier behind a
// walk that re-loaded shared memory every step.  Now:
//   phase 1  128-bit coalesced loads of r, V (done: 32-bit) straight into registers, a whole half
//            tile in flight per thread; delta = (r + gamma*V[t+1]) - V goes to shared memory (V[t+1]
//            of a lane's last step comes from the next lane by shuffle); four ballots of the done
//            flags, bit-interleaved, ARE the ordered episode list;
//   phase 2  one LANE per episode walks backwards over its deltas with 128-bit shared accesses,
//            software-pipelined two groups ahead; the tile's first episode (whose result the previous
//            tile waits for) and its open tail (which waits for the next tile) get warps of their own;
//   phase 3  advantages from shared memory, V again from global (an L2 hit: this CTA read it
//            microseconds ago), returns = adv + V, 128-bit streaming stores, fp64 moments.
// Only the 4 bytes per step that the walk needs live in shared memory (6 144 steps = 26 KiB per CTA).
// HBM traffic stays at the algorithmic 9 B read + 8 B written per step.
#include "g2048_common.cuh"

namespace g2048 {

constexpr int GAE3_THREADS = 256;
constexpr int GAE3_VEC_PER_THREAD = 6;                          // float4 groups per thread
constexpr int GAE3_TILE = GAE3_THREADS * GAE3_VEC_PER_THREAD * 4;  // 6144 steps
constexpr int GAE3_CHUNKS = GAE3_TILE / 32;                     // 192 done masks of 32 steps
constexpr int GAE3_HALF = GAE3_VEC_PER_THREAD / 2;
#ifndef GAE3_MIN_CTAS
#define GAE3_MIN_CTAS 6
#endif

#ifdef G2048_GAE_TIMELINE  // tools/probes/probe_gae3.cu only: per-CTA phase timestamps
__device__ long long g_gae_timeline[16 * 65536];
#define GAE3_STAMP(k) do { if (threadIdx.x == 0 && blockIdx.x < 65536) g_gae_timeline[16 * blockIdx.x + (k)] = clock64(); } while (0)
#define GAE3_STAMP_ANY(k) do { if (blockIdx.x < 65536) g_gae_timeline[16 * blockIdx.x + (k)] = clock64(); } while (0)
#else
#define GAE3_STAMP(k) do { } while (0)
#define GAE3_STAMP_ANY(k) do { } while (0)
#endif

struct Gae3Scratch {
    unsigned int ticket;
    unsigned int pad[3];
};

struct Gae3Smem {
    float g[GAE3_TILE];               // delta -> advantages, in place
    uint32_t mask[GAE3_CHUNKS];       // done bits of each 32-step chunk
    uint32_t pref[GAE3_CHUNKS + 1];   // exclusive prefix of popc(mask)
    double red[4 * (GAE3_THREADS / 32)];
    unsigned int ticket;
};

// position (in the tile) of the k-th done step, k < n_done
__device__ __forceinline__ int gae3_locate(const Gae3Smem& s, int k) {
    int lo = 0, hi = GAE3_CHUNKS - 1;  // largest j with pref[j] <= k
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((int)s.pref[mid] <= k) lo = mid; else hi = mid - 1;
    }
    uint32_t m = s.mask[lo];
    for (int r = k - (int)s.pref[lo]; r > 0; --r) m &= m - 1;  // drop the r lowest set bits
    return 32 * lo + (__ffs((int)m) - 1);
}

// gae = delta + gl * gae backwards over the steps (first_excl, last], in place
__device__ __forceinline__ void gae3_walk(float* __restrict__ sg, int last, int first_excl, float g, float gl) {
    // keep gamma*lambda in a register: left to itself ptxas re-loads it from the constant bank at the top of
    // every trip (LDC, ~30 cycles) and the first multiply of the serial chain waits for it
    asm volatile("" : "+f"(gl));
    int t = last;
    while (t > first_excl && (t & 3) != 3) {  // down to a 16-byte boundary
        g = sg[t] + gl * g;
        sg[t] = g;
        --t;
    }
    // 4-step groups.  Register discipline matters more than instruction count here: the STS.128 of a group
    // holds its source registers until the (conflict-laden, queued) store is dispatched, so neither the
    // serial chain nor the prefetching loads may write those registers soon after.  Inputs (i*, n*) and
    // outputs (o*) therefore live in separate registers, the two input sets swap roles every half trip (no
    // moves), and an output set is rewritten only two groups after its store was issued.
#define GAE3_LD(at) (*reinterpret_cast<const float4*>(&sg[(at)]))
#define GAE3_GROUP(in, out, at)                                  \
    g = in.w + gl * g; out.w = g;                                \
    g = in.z + gl * g; out.z = g;                                \
    g = in.y + gl * g; out.y = g;                                \
    g = in.x + gl * g; out.x = g;                                \
    *reinterpret_cast<float4*>(&sg[(at)]) = out;
    if (t - 4 >= first_excl) {                // steps t-3 .. t are all inside the episode
        float4 i0 = GAE3_LD(t - 3), i1 = i0, n0 = i0, n1 = i0, o0, o1;
        if (t - 8 >= first_excl) i1 = GAE3_LD(t - 7);
        while (true) {
            {   // consume i0, i1; fetch n0, n1 (two and three groups ahead)
                const bool has1 = t - 8 >= first_excl, has2 = t - 12 >= first_excl;
                if (has2) n0 = GAE3_LD(t - 11);
                if (t - 16 >= first_excl) n1 = GAE3_LD(t - 15);
                GAE3_GROUP(i0, o0, t - 3)
                if (!has1) { t -= 4; break; }
                GAE3_GROUP(i1, o1, t - 7)
                t -= 8;
                if (!has2) break;
            }
            {   // consume n0, n1; fetch i0, i1
                const bool has1 = t - 8 >= first_excl, has2 = t - 12 >= first_excl;
                if (has2) i0 = GAE3_LD(t - 11);
                if (t - 16 >= first_excl) i1 = GAE3_LD(t - 15);
                GAE3_GROUP(n0, o0, t - 3)
                if (!has1) { t -= 4; break; }
                GAE3_GROUP(n1, o1, t - 7)
                t -= 8;
                if (!has2) break;
            }
        }
    }
#undef GAE3_GROUP
#undef GAE3_LD
    for (; t > first_excl; --t) {
        g = sg[t] + gl * g;
        sg[t] = g;
    }
}

// bits 0..7 of b -> bit positions 0, 4, 8, ..., 28
__device__ __forceinline__ uint32_t spread_bits4(uint32_t b) {
    uint32_t x = b & 0xFFu;
    x = (x | (x << 12)) & 0x000F000Fu;
    x = (x | (x << 6)) & 0x03030303u;
    x = (x | (x << 3)) & 0x11111111u;
    return x;
}

template <bool ALIGNED>
__device__ __forceinline__ float4 gae3_load4(const float* __restrict__ p, int64_t gi, int valid) {
    if (ALIGNED && valid == 4) return __ldg(reinterpret_cast<const float4*>(p + gi));
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid > 0) o.x = __ldg(p + gi);
    if (valid > 1) o.y = __ldg(p + gi + 1);
    if (valid > 2) o.z = __ldg(p + gi + 2);
    if (valid > 3) o.w = __ldg(p + gi + 3);
    return o;
}

template <bool ALIGNED>
__device__ __forceinline__ uint32_t gae3_load_done4(const uint8_t* __restrict__ p, int64_t gi, int valid) {
    uint32_t w = 0;
    if (ALIGNED && valid == 4) {
        w = __ldg(reinterpret_cast<const uint32_t*>(p + gi));
    } else {
        if (valid > 0) w |= (uint32_t)__ldg(p + gi);
        if (valid > 1) w |= (uint32_t)__ldg(p + gi + 1) << 8;
        if (valid > 2) w |= (uint32_t)__ldg(p + gi + 2) << 16;
        if (valid > 3) w |= (uint32_t)__ldg(p + gi + 3) << 24;
    }
    return w;
}

template <bool ALIGNED>
__global__ void __launch_bounds__(GAE3_THREADS, GAE3_MIN_CTAS)
gae_flat3_kernel(const float* __restrict__ rewards, const float* __restrict__ values, const uint8_t* __restrict__ dones,
                 int64_t n, int64_t n_tiles, float gamma, float gamma_lambda, float* __restrict__ adv,
                 float* __restrict__ ret, Gae3Scratch* scratch, double* __restrict__ moments) {
    __shared__ __align__(16) Gae3Smem s;
    volatile unsigned int* flags = (volatile unsigned int*)(scratch + 1);
    volatile float* heads = (volatile float*)((unsigned int*)(scratch + 1) + n_tiles);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    GAE3_STAMP(0);
    if (tid == 0) s.ticket = atomicAdd(&scratch->ticket, 1u);
    __syncthreads();
    const int64_t tile = n_tiles - 1 - (int64_t)s.ticket;  // tiles are taken from the END of the buffer
    const int64_t lo = tile * GAE3_TILE;
    const int len = (int)min((int64_t)GAE3_TILE, n - lo);
    GAE3_STAMP(1);

    // ---- phase 1: delta into shared memory, done bits into ordered masks --------------------------------
    // thread t owns the 4 consecutive steps 4*(q*256 + t) .. +3 of group q; a warp covers 128 steps
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float4 r[GAE3_HALF], v[GAE3_HALF];
        uint32_t d[GAE3_HALF];
        float vnext[GAE3_HALF];
#pragma unroll
        for (int k = 0; k < GAE3_HALF; ++k) {
            const int i = 4 * ((h * GAE3_HALF + k) * GAE3_THREADS + tid);
            const int valid = max(0, min(4, len - i));
            r[k] = gae3_load4<ALIGNED>(rewards, lo + i, valid);
            v[k] = gae3_load4<ALIGNED>(values, lo + i, valid);
            d[k] = gae3_load_done4<ALIGNED>(dones, lo + i, valid);
            // V of the step after this lane's four: the next lane has it, except for lane 31
            vnext[k] = (lane == 31 && lo + i + 4 < n && i + 4 <= len + 3) ? __ldg(values + lo + i + 4) : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < GAE3_HALF; ++k) {
            const int q = h * GAE3_HALF + k;
            const int i = 4 * (q * GAE3_THREADS + tid);
            const float from_next_lane = __shfl_down_sync(0xFFFFFFFFu, v[k].x, 1);
            const float v4 = (lane == 31) ? vnext[k] : from_next_lane;  // 0 past the end of the buffer
            const bool d0 = (d[k] & 0xFFu) != 0, d1 = (d[k] & 0xFF00u) != 0, d2 = (d[k] & 0xFF0000u) != 0,
                       d3 = (d[k] & 0xFF000000u) != 0;
            float4 delta;
            delta.x = (r[k].x + gamma * (d0 ? 0.0f : v[k].y)) - v[k].x;
            delta.y = (r[k].y + gamma * (d1 ? 0.0f : v[k].z)) - v[k].y;
            delta.z = (r[k].z + gamma * (d2 ? 0.0f : v[k].w)) - v[k].z;
            delta.w = (r[k].w + gamma * (d3 ? 0.0f : v4)) - v[k].w;
            *reinterpret_cast<float4*>(&s.g[i]) = delta;
            const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, d0), b1 = __ballot_sync(0xFFFFFFFFu, d1),
                           b2 = __ballot_sync(0xFFFFFFFFu, d2), b3 = __ballot_sync(0xFFFFFFFFu, d3);
            if (lane < 4) {  // lane c assembles the mask of the warp's c-th 32-step chunk (lanes 8c .. 8c+7)
                const int sh = 8 * lane;
                const uint32_t m = spread_bits4(b0 >> sh) | (spread_bits4(b1 >> sh) << 1) |
                                   (spread_bits4(b2 >> sh) << 2) | (spread_bits4(b3 >> sh) << 3);
                s.mask[(q * GAE3_THREADS + 32 * warp) / 8 + lane] = m;
            }
        }
    }
    __syncthreads();
    GAE3_STAMP(2);
    // exclusive prefix over the 192 chunk counts (one warp, 6 chunks per lane)
    if (warp == 0) {
        uint32_t c[GAE3_CHUNKS / 32];
        uint32_t sum = 0;
#pragma unroll
        for (int q = 0; q < GAE3_CHUNKS / 32; ++q) {
            c[q] = (uint32_t)__popc(s.mask[lane * (GAE3_CHUNKS / 32) + q]);
            sum += c[q];
        }
        uint32_t incl = sum;
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, off);
            if (lane >= off) incl += y;
        }
        uint32_t run = incl - sum;
#pragma unroll
        for (int q = 0; q < GAE3_CHUNKS / 32; ++q) {
            s.pref[lane * (GAE3_CHUNKS / 32) + q] = run;
            run += c[q];
        }
        if (lane == 31) s.pref[GAE3_CHUNKS] = incl;
    }
    __syncthreads();
    const int n_done = (int)s.pref[GAE3_CHUNKS];
    GAE3_STAMP(3);

    // ---- phase 2: one lane per episode ---------------------------------------------------------------------
    // Warp w issues on scheduler w % 4.  The walking warps of the CTAs that share an SM must not pile up
    // on one scheduler (timeline probe: with fixed roles the five "warp 0" walkers of an SM shared
    // scheduler 0 and the walk ran at 35 cycles per step), so the roles rotate with the ticket.
    const int rot = (int)(s.ticket & 3u);
    const int role = (warp < 4) ? ((warp - rot) & 3) : 4 + ((warp - rot) & 3);  // 0-3, 6-7 walk; 4 first episode; 5 tail
    if (role == 4) {
        // the first episode of the tile alone in its warp: the previous tile is waiting for its result
        if (lane == 0 && n_done > 0) {
#ifndef G2048_GAE3_SKIP_SIDE_WALKS  // timing experiment only (tools/probes/probe_gae3.cu)
            gae3_walk(s.g, gae3_locate(s, 0), -1, 0.0f, gamma_lambda);
#endif
            heads[tile] = s.g[0];
            __threadfence();
            flags[tile] = 1u;
            GAE3_STAMP_ANY(8);   // first episode published
        }
    } else if (role == 5) {
        // the steps after the tile's last done belong to an episode that ends in a later tile
        if (lane == 31) {
            const int first_excl = n_done ? gae3_locate(s, n_done - 1) : -1;
            if (first_excl < len - 1) {
                float carry = 0.0f;
                if (lo + len < n) {
                    while (flags[tile + 1] == 0u) __nanosleep(20);
                    __threadfence();
                    carry = heads[tile + 1];
                }
                GAE3_STAMP_ANY(9);   // look-back satisfied
#ifndef G2048_GAE3_SKIP_SIDE_WALKS
                gae3_walk(s.g, len - 1, first_excl, carry, gamma_lambda);
#endif
                GAE3_STAMP_ANY(10);  // tail walked
            }
            if (n_done == 0) {
                heads[tile] = s.g[0];
                __threadfence();
                flags[tile] = 1u;
            }
        }
    } else {
        // episodes 1.. : the first 32 on the role-0 warp, the next 32 on role 1, ... (six walking roles)
        const int slot = role < 4 ? role : role - 2;
        for (int e = 1 + 32 * slot + lane; e < n_done; e += 6 * 32) {
            gae3_walk(s.g, gae3_locate(s, e), gae3_locate(s, e - 1), 0.0f, gamma_lambda);
        }
        if (slot == 0 && lane == 0) GAE3_STAMP_ANY(11);  // the first walking warp finished its episodes
    }
    __syncthreads();
    GAE3_STAMP(4);

    // ---- phase 3: returns, stores, moments -------------------------------------------------------------------
    double m[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float4 v[GAE3_HALF];
#pragma unroll
        for (int k = 0; k < GAE3_HALF; ++k) {
            const int i = 4 * ((h * GAE3_HALF + k) * GAE3_THREADS + tid);
            v[k] = gae3_load4<ALIGNED>(values, lo + i, max(0, min(4, len - i)));  // L2 hit
        }
#pragma unroll
        for (int k = 0; k < GAE3_HALF; ++k) {
            const int i = 4 * ((h * GAE3_HALF + k) * GAE3_THREADS + tid);
            const int valid = max(0, min(4, len - i));
            if (valid > 0) {
                const float4 a = *reinterpret_cast<const float4*>(&s.g[i]);
                const float4 rt = make_float4(a.x + v[k].x, a.y + v[k].y, a.z + v[k].z, a.w + v[k].w);
                if (ALIGNED && valid == 4) {
                    __stcs(reinterpret_cast<float4*>(adv + lo + i), a);
                    __stcs(reinterpret_cast<float4*>(ret + lo + i), rt);
                } else {
                    const float aa[4] = {a.x, a.y, a.z, a.w}, rr[4] = {rt.x, rt.y, rt.z, rt.w};
                    for (int j = 0; j < valid; ++j) {
                        adv[lo + i + j] = aa[j];
                        ret[lo + i + j] = rr[j];
                    }
                }
                const float aa[4] = {a.x, a.y, a.z, a.w}, rr[4] = {rt.x, rt.y, rt.z, rt.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (j < valid) {
                        m[0] += (double)aa[j];
                        m[1] += (double)aa[j] * (double)aa[j];
                        m[2] += (double)rr[j];
                        m[3] += (double)rr[j] * (double)rr[j];
                    }
                }
            }
        }
    }
    GAE3_STAMP(5);
    if (moments) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double x = m[k];
            for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xFFFFFFFFu, x, off);
            if (lane == 0) s.red[k * (GAE3_THREADS / 32) + warp] = x;
        }
        __syncthreads();
        if (tid < 4) {
            double t = 0.0;
            for (int w = 0; w < GAE3_THREADS / 32; ++w) t += s.red[tid * (GAE3_THREADS / 32) + w];
            atomicAdd(&moments[1 + tid], t);
        }
        if (tid == 4) atomicAdd(&moments[0], (double)len);
    }
    GAE3_STAMP(6);
}

}  // namespace g2048

using namespace g2048;

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

extern "C" int g2048_gae_flat(const float* d_rewards, const float* d_values, const uint8_t* d_dones, int64_t n,
                              double gamma, double lambda_gae, float* d_adv, float* d_ret, void* d_scan_state,
                              double* d_moments, void* stream) {
    G2048_REQUIRE(n >= 0, "gae_flat: n");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_rewards && d_values && d_dones && d_adv && d_ret && d_scan_state, "gae_flat: pointers");
    const int64_t n_tiles = (n + GAE3_TILE - 1) / GAE3_TILE;
    // 128-bit accesses need 16-byte aligned float arrays and a 4-byte aligned done array
    const bool aligned = aligned16(d_rewards) && aligned16(d_values) && aligned16(d_adv) && aligned16(d_ret) &&
                         ((uintptr_t)d_dones & 3u) == 0;
    static bool configured = false;
    if (!configured) {  // several CTAs of ~26 KiB per SM need the shared-memory-heavy L1 split
        cudaFuncSetAttribute(gae_flat3_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(gae_flat3_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        configured = true;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (aligned) {
        gae_flat3_kernel<true><<<(unsigned)n_tiles, GAE3_THREADS, 0, st>>>(
            d_rewards, d_values, d_dones, n, n_tiles, (float)gamma, (float)(gamma * lambda_gae), d_adv, d_ret,
            (Gae3Scratch*)d_scan_state, d_moments);
    } else {
        gae_flat3_kernel<false><<<(unsigned)n_tiles, GAE3_THREADS, 0, st>>>(
            d_rewards, d_values, d_dones, n, n_tiles, (float)gamma, (float)(gamma * lambda_gae), d_adv, d_ret,
            (Gae3Scratch*)d_scan_state, d_moments);
    }
    G2048_CHECK_LAUNCH("gae_flat");
    return G2048_OK;
}
