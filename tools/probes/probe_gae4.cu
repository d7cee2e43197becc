// Phase timeline of the flat GAE kernel: builds g2048_gae4.cu with -DG2048_GAE_TIMELINE and prints, per phase,
// the mean / p50 / p95 duration in SM cycles over all CTAs, plus the kernel time.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define G2048_GAE4_NO_DISPATCH 1
#include "../../2048-ppo-agent_b200/csrc/g2048_gae4.cu"
namespace g2048 { thread_local char g_last_error[512] = ""; int sm_count() { int n = 0; cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, 0); return n; } }
extern "C" int64_t g2048_gae_flat_scratch_bytes(int64_t n) { return 16 + ((n + 1023) / 1024) * 8; }
extern "C" const char* g2048_last_error(void) { return g2048::g_last_error; }

int main(int argc, char** argv) {
    const int64_t n = 1ll << (argc > 2 ? atoi(argv[2]) : 26);
    const double rate = argc > 1 ? atof(argv[1]) : 1.0 / 300;
    std::vector<float> r(n), v(n); std::vector<uint8_t> d(n);
    uint64_t s = 88172645463325252ull;
    for (int64_t i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; r[i] = (s & 255) * 0.25f; v[i] = ((s >> 8) & 1023) / 512.0f - 1; d[i] = ((s >> 20) % 1000000) < rate * 1e6; }
    float *dr, *dv, *da, *dt; uint8_t* dd; void* ds; double* dm;
    cudaMalloc(&dr, n * 4); cudaMalloc(&dv, n * 4); cudaMalloc(&da, n * 4); cudaMalloc(&dt, n * 4); cudaMalloc(&dd, n);
    const int64_t sb = g2048_gae_flat_scratch_bytes(n);
    cudaMalloc(&ds, sb); cudaMalloc(&dm, 48);
    cudaMemcpy(dr, r.data(), n * 4, cudaMemcpyHostToDevice); cudaMemcpy(dv, v.data(), n * 4, cudaMemcpyHostToDevice); cudaMemcpy(dd, d.data(), n, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
        cudaMemset(ds, 0, sb); cudaMemset(dm, 0, 48);
        cudaEventRecord(e0);
        int rc = g2048_gae_flat_pipelined(dr, dv, dd, n, 0.99, 0.95, da, dt, ds, dm, nullptr);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        if (rc) { printf("rc %d %s\n", rc, g2048_last_error()); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const int64_t tiles = (n + g2048::G4_TILE - 1) / g2048::G4_TILE;
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, g2048::gae_flat4_kernel<true>, g2048::G4_THREADS, sizeof(g2048::G4Smem));
    printf("n=2^%d rate %.5f: %.1f us, %.0f GB/s, %lld tiles, %d CTAs/SM by occupancy API\n", (int)(argc > 2 ? atoi(argv[2]) : 26), rate, best * 1e3, n * 17.0 / best / 1e6, (long long)tiles, per_sm);
    return 0;
}
