// How fast is the per-episode GAE walk (g2048_gae3.cu: gae3_walk) by itself?  One CTA per SM, W walking warps with
// L active lanes each, every lane walks its own 1024-step segment of shared memory.  Prints cycles per step.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../2048-ppo-agent_b200/csrc/g2048_gae3.cu"
namespace g2048 { thread_local char g_last_error[512] = ""; int sm_count() { return 148; } }
extern "C" int64_t g2048_gae_flat_scratch_bytes(int64_t n) { return 16 + ((n + 1023) / 1024) * 8; }
extern "C" const char* g2048_last_error(void) { return g2048::g_last_error; }

__global__ void walk_probe(int warps, int lanes, int steps, int stride, long long* out, float* sink) {
    extern __shared__ __align__(16) float sg[];
    for (int i = threadIdx.x; i < 48 * 1024 / 4 * 4; i += blockDim.x) sg[i] = 0.001f * (i & 1023);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long t0 = clock64();
    float g = 0.f;
    if (warp < warps && lane < lanes) {
        const int base = ((warp * 32 + lane) * stride) % (48 * 1024 - steps - 8);
        g2048::gae3_walk(sg, base + steps - 1, base - 1, 0.0f, 0.9405f);
        g = sg[base];
    }
    long long t1 = clock64();
    if (lane == 0 && warp < warps && blockIdx.x == 0) out[warp] = t1 - t0;
    if (g == 12345.f) sink[0] = g;
}

int main() {
    long long* d_out; float* d_sink;
    cudaMalloc(&d_out, 64 * 8); cudaMalloc(&d_sink, 4);
    cudaFuncSetAttribute(walk_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024);
    const int steps = 1024;
    for (int stride : {1031, 1024}) {
        for (int warps : {1, 2, 4, 8}) {
            for (int lanes : {1, 8, 20, 32}) {
                walk_probe<<<148, 256, 192 * 1024>>>(warps, lanes, steps, stride, d_out, d_sink);
                cudaDeviceSynchronize();
                walk_probe<<<148, 256, 192 * 1024>>>(warps, lanes, steps, stride, d_out, d_sink);
                long long h[8];
                cudaMemcpy(h, d_out, 64, cudaMemcpyDeviceToHost);
                long long mx = 0; for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
                printf("stride %4d warps %d lanes %2d: %6.1f cycles/step\n", stride, warps, lanes, (double)mx / steps);
            }
        }
    }
    return 0;
}
