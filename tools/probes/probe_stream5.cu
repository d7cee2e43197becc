// What does the GAE traffic pattern itself allow?  3 read streams (f32, f32, u8) + 2 write streams (f32, f32), no recurrence:
// adv = r + v * d, ret = r - v, grid-stride with 128-bit accesses.  Compare with the GAE kernels' GB/s at the same size.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o probe_stream5 probe_stream5.cu && ./probe_stream5
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) stream5(const float4* __restrict__ r, const float4* __restrict__ v, const uchar4* __restrict__ d,
                                               float4* __restrict__ a, float4* __restrict__ q, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 x = r[i], y = v[i];
        const uchar4 m = d[i];
        a[i] = make_float4(x.x + y.x * m.x, x.y + y.y * m.y, x.z + y.z * m.z, x.w + y.w * m.w);
        q[i] = make_float4(x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w);
    }
}
// the same traffic, but every CTA walks its own contiguous range of the buffer backwards in 2 048-element tiles (the scan
// kernel's mapping) instead of the grid-stride wave front
__global__ void __launch_bounds__(256) stream5_ranges(const float4* __restrict__ r, const float4* __restrict__ v, const uchar4* __restrict__ d,
                                                      float4* __restrict__ a, float4* __restrict__ q, int64_t n4, int64_t tiles_per_range) {
    const int64_t n_tiles = n4 / 512;  // 512 float4 = 2 048 elements per tile
    const int64_t lo = (int64_t)blockIdx.x * tiles_per_range, hi = min(n_tiles, lo + tiles_per_range);
    for (int64_t t = hi - 1; t >= lo; --t) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int64_t i = t * 512 + k * 256 + threadIdx.x;
            const float4 x = r[i], y = v[i];
            const uchar4 m = d[i];
            a[i] = make_float4(x.x + y.x * m.x, x.y + y.y * m.y, x.z + y.z * m.z, x.w + y.w * m.w);
            q[i] = make_float4(x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w);
        }
    }
}

__global__ void __launch_bounds__(256) copy2(const float4* __restrict__ r, float4* __restrict__ a, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) a[i] = r[i];
}

int main() {
    const int64_t n = 1ll << 26, n4 = n / 4;
    float *r, *v, *a, *q;
    uint8_t* d;
    void* flush;
    cudaMalloc(&r, n * 4); cudaMalloc(&v, n * 4); cudaMalloc(&a, n * 4); cudaMalloc(&q, n * 4); cudaMalloc(&d, n);
    cudaMalloc(&flush, 512 << 20);
    cudaMemset(r, 0, n * 4); cudaMemset(v, 0, n * 4); cudaMemset(d, 0, n);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mult : {8, 16, 32}) {
        float best5 = 1e9f, bestc = 1e9f;
        for (int rep = 0; rep < 6; ++rep) {
            cudaMemset(flush, rep, 512 << 20);
            cudaEventRecord(e0);
            stream5<<<sms * mult, 256>>>((const float4*)r, (const float4*)v, (const uchar4*)d, (float4*)a, (float4*)q, n4);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best5) best5 = ms;
            cudaMemset(flush, rep, 512 << 20);
            cudaEventRecord(e0);
            copy2<<<sms * mult, 256>>>((const float4*)r, (float4*)a, n4);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < bestc) bestc = ms;
        }
        printf("grid %d x SMs: stream5 (17 B/elem) %.1f us = %.2f TB/s   copy (8 B/elem) %.1f us = %.2f TB/s\n", mult, best5 * 1e3,
               n * 17.0 / best5 / 1e9, bestc * 1e3, n * 8.0 / bestc / 1e9);
    }
    for (int per_sm : {3, 4, 8}) {
        const int grid = sms * per_sm;
        const int64_t tpr = (n4 / 512 + grid - 1) / grid;
        float best = 1e9f;
        for (int rep = 0; rep < 6; ++rep) {
            cudaMemset(flush, rep, 512 << 20);
            cudaEventRecord(e0);
            stream5_ranges<<<grid, 256>>>((const float4*)r, (const float4*)v, (const uchar4*)d, (float4*)a, (float4*)q, n4, tpr);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        printf("ranges, %d CTAs per SM: stream5 %.1f us = %.2f TB/s\n", per_sm, best * 1e3, n * 17.0 / best / 1e9);
    }
    return 0;
}
