// Host <-> device copies for the `*_host` entry points.  Callers hand in ordinary (pageable) arrays -- numpy, malloc --
// and a plain cudaMemcpy of pageable memory is staged by the driver, synchronously, at a few GB/s.  StagedCopier keeps
// two pinned 16 MiB buffers per calling thread and pipelines the transfer through them (host memcpy of chunk k+1 while
// chunk k is on the bus).  The host-side memcpy of a chunk is split over a few threads: one core moves ~10 GB/s, a
// PCIe 5 x16 link five times that, so with a single copying thread the 256 MiB of per-env results of a 2^24-env batch
// cost 25 ms behind a 70 ms kernel.  Memory the caller already pinned (cudaHostAlloc / cudaHostRegister) is copied directly.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <thread>

#include "g2048_common.cuh"

namespace g2048 {

struct StagedCopier {
    static constexpr size_t STAGE = 16u << 20;
    uint8_t* pin[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};

    int init() {
        for (int k = 0; k < 2; ++k) {
            int rc = check_cuda(cudaMallocHost((void**)&pin[k], STAGE), "host copy: pinned staging");
            if (!rc) rc = check_cuda(cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming), "host copy: event");
            if (rc) return rc;
        }
        return G2048_OK;
    }

    // give the staging buffers and events back (errors ignored: at thread or process exit the context may be gone)
    void release() {
        for (int k = 0; k < 2; ++k) {
            if (pin[k]) cudaFreeHost(pin[k]);
            if (ev[k]) cudaEventDestroy(ev[k]);
            pin[k] = nullptr;
            ev[k] = nullptr;
        }
        cudaGetLastError();
    }

    // memcpy of one staged chunk; pieces of at least 1 MiB on up to G2048_HOSTCOPY_THREADS (default 4) threads
    static void copy_chunk(void* dst, const void* src, size_t bytes) {
        static const int max_threads = [] {
            const char* e = getenv("G2048_HOSTCOPY_THREADS");
            int v = e ? atoi(e) : 4;
            const int hw = (int)std::thread::hardware_concurrency();
            if (hw > 0 && v > hw) v = hw;
            return v < 1 ? 1 : (v > 16 ? 16 : v);
        }();
        int parts = (int)(bytes >> 20);
        if (parts > max_threads) parts = max_threads;
        if (parts <= 1) {
            memcpy(dst, src, bytes);
            return;
        }
        const size_t piece = ((bytes / parts) + 4095) & ~(size_t)4095;
        std::thread helpers[16];
        for (int k = 1; k < parts; ++k) {
            const size_t off = (size_t)k * piece;
            const size_t len = off >= bytes ? 0 : (bytes - off < piece ? bytes - off : piece);
            helpers[k] = std::thread([=] { if (len) memcpy((char*)dst + off, (const char*)src + off, len); });
        }
        memcpy(dst, src, piece < bytes ? piece : bytes);
        for (int k = 1; k < parts; ++k) helpers[k].join();
    }

    static bool is_pinned(const void* p) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
            cudaGetLastError();  // older runtimes report unregistered memory as an error: clear it
            return false;
        }
        return attr.type == cudaMemoryTypeHost;
    }

    int h2d(void* d_dst, const void* h_src, size_t bytes, cudaStream_t st) {
        if (is_pinned(h_src)) return check_cuda(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st), "host copy: h2d");
        int k = 0;
        for (size_t off = 0; off < bytes; k ^= 1) {
            const size_t c = bytes - off < STAGE ? bytes - off : STAGE;
            int rc = check_cuda(cudaEventSynchronize(ev[k]), "host copy: staging");  // the copy that last used this buffer
            if (rc) return rc;
            copy_chunk(pin[k], (const char*)h_src + off, c);
            rc = check_cuda(cudaMemcpyAsync((char*)d_dst + off, pin[k], c, cudaMemcpyHostToDevice, st), "host copy: h2d");
            if (!rc) rc = check_cuda(cudaEventRecord(ev[k], st), "host copy: staging");
            if (rc) return rc;
            off += c;
        }
        return G2048_OK;
    }

    // On return the data of a PAGEABLE destination is in place; a pinned destination is only enqueued (synchronise `st`).
    int d2h(void* h_dst, const void* d_src, size_t bytes, cudaStream_t st) {
        if (is_pinned(h_dst)) return check_cuda(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, st), "host copy: d2h");
        int k = 0;
        size_t prev_off = 0, prev_c = 0;
        for (size_t off = 0; off < bytes || prev_c; k ^= 1) {
            size_t c = 0;
            if (off < bytes) {  // chunk i onto the bus ...
                c = bytes - off < STAGE ? bytes - off : STAGE;
                int rc = check_cuda(cudaMemcpyAsync(pin[k], (const char*)d_src + off, c, cudaMemcpyDeviceToHost, st), "host copy: d2h");
                if (!rc) rc = check_cuda(cudaEventRecord(ev[k], st), "host copy: staging");
                if (rc) return rc;
            }
            if (prev_c) {  // ... while chunk i-1 goes from its staging buffer to the caller's array
                const int rc = check_cuda(cudaEventSynchronize(ev[k ^ 1]), "host copy: staging");
                if (rc) return rc;
                copy_chunk((char*)h_dst + prev_off, pin[k ^ 1], prev_c);
            }
            prev_off = off;
            prev_c = c;
            off += c;
        }
        return G2048_OK;
    }
};

}  // namespace g2048
