// GAE kernel experiments: times g2048_gae_flat built with -DG2048_GAE_EXPERIMENT=<mask>
//   1 = no serial walk (adv = delta), 2 = no moments, 4 = no look-back wait
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../2048-ppo-agent_b200/csrc/g2048_gae.cu"
namespace g2048 { thread_local char g_last_error[512] = ""; int sm_count() { int n = 0; cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, 0); return n; } }

int main(int argc, char** argv) {
    const int64_t n = 1ll << 26;
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
    for (int rep = 0; rep < 5; ++rep) {
        cudaMemset(ds, 0, sb); cudaMemset(dm, 0, 48);
        cudaEventRecord(e0);
        int rc = g2048_gae_flat_v1(dr, dv, dd, n, 0.99, 0.95, da, dt, ds, (G2048_GAE_EXPERIMENT & 2) ? nullptr : dm, nullptr);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        if (rc) { printf("rc %d %s\n", rc, g2048_last_error ? "" : ""); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("experiment mask %d, done rate %.5f: %.1f us, %.0f GB/s\n", G2048_GAE_EXPERIMENT, rate, best * 1e3, n * 17.0 / best / 1e6);
    return 0;
}
