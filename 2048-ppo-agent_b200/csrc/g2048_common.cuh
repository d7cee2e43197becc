// Shared host-side helpers of libg2048.so: error reporting and launch geometry.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "../../include/g2048.h"

namespace g2048 {

// last error text for g2048_last_error(); thread_local so concurrent callers do not race
extern thread_local char g_last_error[512];

inline int fail_arg(const char* what) {
    snprintf(g_last_error, sizeof(g_last_error), "g2048: invalid argument: %s", what);
    return G2048_ERR_INVALID;
}

inline int check_cuda(cudaError_t e, const char* where) {
    if (e == cudaSuccess) return G2048_OK;
    snprintf(g_last_error, sizeof(g_last_error), "g2048: CUDA error in %s: %s", where, cudaGetErrorString(e));
    return (int)e;
}

#define G2048_CHECK_LAUNCH(where) \
    do { int _rc = ::g2048::check_cuda(cudaGetLastError(), where); if (_rc) return _rc; } while (0)

#define G2048_REQUIRE(cond, what) \
    do { if (!(cond)) return ::g2048::fail_arg(what); } while (0)

struct DeviceInfo {
    int sm_count;
    int device;
};

// SM count of the current device (cached per device)
int sm_count();

// One-time per-device set-up (cudaFuncSetAttribute is per device): `flags` is a static bool[64] of the call site.
// Returns a pointer to the current device's flag, or nullptr when there is no usable device.
inline bool* device_once_flag(bool (&flags)[64]) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    return &flags[dev];
}

// Race check by perturbation (compute-sanitizer's racecheck only sees shared memory; the protocols worth worrying about
// here go through GLOBAL memory: tile look-back flags, work queues, the step ticket).  A build with -DG2048_RACE_JITTER
// (make jitter -> tests/legacy/libg2048_jitter.so) delays threads by a pseudo-random 0.2 - 1.2 us at every such
// hand-over, one call in eight; tests/test_race_jitter_gpu.py requires bit-identical results from it.
#ifdef __CUDACC__
__device__ __forceinline__ void race_jitter() {
#ifdef G2048_RACE_JITTER
    const unsigned c = (unsigned)clock64() * 2654435761u + threadIdx.x * 40503u + blockIdx.x * 977u;
    if (((c >> 9) & 7u) == 0u) __nanosleep(200u + (c & 1023u));
#endif
}
#endif

inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace g2048
