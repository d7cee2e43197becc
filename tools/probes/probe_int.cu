// Integer-pipe probes: which instruction mix for a Threefry round issues fastest on sm_100a?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o probe_int probe_int.cu && ./probe_int
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t rotl_shf(uint32_t x, int r) { return __funnelshift_l(x, x, r); }

// rotate through the FMA pipe: 64-bit product x * 2^r = {x >> (32-r), x << r}
__device__ __forceinline__ void mulwide(uint32_t x, uint32_t pow2, uint32_t& lo, uint32_t& hi) {
    asm volatile("{ .reg .u64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t; }" : "=r"(lo), "=r"(hi) : "r"(x), "r"(pow2));
}
__device__ __forceinline__ uint32_t add_mad(uint32_t a, uint32_t b) {
    uint32_t d;
    asm volatile("mad.lo.u32 %0, %1, 1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

template <int V>
__device__ __forceinline__ void round_v(uint32_t& a, uint32_t& b, int r, int parity) {
    if (V == 0) { a += b; b = rotl_shf(b, r); b ^= a; }
    if (V == 1) { a += b; uint32_t lo, hi; mulwide(b, 1u << r, lo, hi); b = (lo | hi) ^ a; }
    if (V == 2) { a = add_mad(a, b); uint32_t lo, hi; mulwide(b, 1u << r, lo, hi); b = (lo | hi) ^ a; }
    if (V == 3) { if (parity) a = add_mad(a, b); else a += b; uint32_t lo, hi; mulwide(b, 1u << r, lo, hi); b = (lo | hi) ^ a; }
    if (V == 4) { a = add_mad(a, b); b = rotl_shf(b, r); b ^= a; }
    if (V == 6) { a = add_mad(a, b); if (parity == 3) { uint32_t lo, hi; mulwide(b, 1u << r, lo, hi); b = (lo | hi) ^ a; } else { b = rotl_shf(b, r); b ^= a; } }  // every 4th rotate through the FMA pipe
    if (V == 7) { a = add_mad(a, b); if (parity == 3) { uint32_t lo, hi; mulwide(b, 1u << r, lo, hi); b = add_mad(lo, hi) ^ a; } else { b = rotl_shf(b, r); b ^= a; } }  // same, halves joined by an add on the FMA pipe
    if (V == 5) { if (parity) { a += b; uint32_t lo, hi; mulwide(b, 1u << r, lo, hi); b = (lo | hi) ^ a; } else { a = add_mad(a, b); b = rotl_shf(b, r); b ^= a; } }
}

template <int V, int CH>
__global__ void probe(int iters, uint32_t* sink) {
    uint32_t a[CH], b[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) { a[c] = threadIdx.x + 17 * c; b[c] = blockIdx.x * 31 + c + 1; }
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int rs[4] = {13, 15, 26, 6};
#pragma unroll
            for (int c = 0; c < CH; ++c) round_v<V>(a[c], b[c], rs[k], V >= 6 ? k : (k & 1));
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) x ^= a[c] ^ b[c];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

template <int V, int CH>
void run(const char* name, int blocks) {
    uint32_t* sink;
    cudaMalloc(&sink, (size_t)blocks * 256 * 4);
    const int iters = 20000;
    probe<V, CH><<<blocks, 256>>>(iters, sink);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        probe<V, CH><<<blocks, 256>>>(iters, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double rounds = (double)blocks * 256 * iters * 4 * CH;
    printf("%-44s CH=%d  %7.3f ms  %7.2f T rounds/s  (%.2f T instr/s at 3 instr/round)\n", name, CH, best, rounds / best / 1e9,
           3 * rounds / best / 1e9);
    cudaFree(sink);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8;
    run<0, 4>("V0 add + SHF + LOP3 (C code)", blocks);
    run<1, 4>("V1 add + IMAD.WIDE rot + LOP3", blocks);
    run<2, 4>("V2 mad-add + IMAD.WIDE rot + LOP3", blocks);
    run<3, 4>("V3 alternating add/mad + IMAD.WIDE + LOP3", blocks);
    run<4, 4>("V4 mad-add + SHF + LOP3", blocks);
    run<5, 4>("V5 alternate (add,WIDE) / (mad,SHF)", blocks);
    run<6, 4>("V6 mad-add; SHF x3 + IMAD.WIDE x1 per 4 rounds", blocks);
    run<7, 4>("V7 same, halves joined by mad-add", blocks);
    run<0, 2>("V0", blocks); run<3, 2>("V3", blocks); run<5, 2>("V5", blocks);
    run<0, 8>("V0", blocks); run<3, 8>("V3", blocks); run<5, 8>("V5", blocks);
    return 0;
}
