/* lbad_common.cuh — error plumbing, PTX wrappers (mbarrier + 1-D bulk TMA copy) and launch timing shared by the .cu files. */
#ifndef LBAD_COMMON_CUH
#define LBAD_COMMON_CUH
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
#include "lbad_cuda.h"

namespace lbad {

void set_error(const char* fmt, ...);

#define LBAD_CUDA_TRY(expr)                                                                         \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            ::lbad::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return (_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver) ? LBAD_ERR_NODEVICE : LBAD_ERR_CUDA; \
        }                                                                                           \
    } while (0)

/* Makes `device` current for the scope of an entry point and puts the caller's device back on the way out: a drop-in library
 * must not leave the host application on another GPU than the one it had selected. */
struct DeviceScope {
    int prev = -1; cudaError_t status = cudaSuccess;
    explicit DeviceScope(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        if (prev != device) status = cudaSetDevice(device);
        else prev = -1;                                     /* nothing to restore */
    }
    DeviceScope(const DeviceScope&) = delete;
    DeviceScope& operator=(const DeviceScope&) = delete;
    ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};
#define LBAD_ON_DEVICE(dev) ::lbad::DeviceScope _device_scope(dev); LBAD_CUDA_TRY(_device_scope.status)

/* cudaMalloc with scope lifetime, so that an early error return does not leak */
template <class T>
struct DevBuf {
    T* p = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, (n ? n : 1) * sizeof(T)); }
    ~DevBuf() { if (p) cudaFree(p); }
    operator T*() const { return p; }
};

/* A half-built object on the error paths of a create function: destroyed on scope exit unless released */
template <class T>
struct Guard {
    T* p; void (*destroy)(T*);
    Guard(T* p_, void (*d)(T*)) : p(p_), destroy(d) {}
    Guard(const Guard&) = delete;
    Guard& operator=(const Guard&) = delete;
    ~Guard() { if (p) destroy(p); }
    T* release() { T* r = p; p = nullptr; return r; }
};

/* Device time of selected launches, measured with CUDA events on the launching stream. */
struct LaunchTimer {
    bool enabled = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
    void begin(cudaStream_t s) {
        if (!enabled) return;
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, s); events.push_back({a, b});
    }
    void end(cudaStream_t s) { if (enabled && !events.empty()) cudaEventRecord(events.back().second, s); }
    uint32_t collect(double* total_ms, bool reset) {
        double t = 0; uint32_t n = 0;
        for (auto& e : events) {
            if (cudaEventSynchronize(e.second) != cudaSuccess) continue;
            float ms = 0; if (cudaEventElapsedTime(&ms, e.first, e.second) == cudaSuccess) { t += ms; n++; }
        }
        if (total_ms) *total_ms = t;
        if (reset) clear();
        return n;
    }
    void clear() { for (auto& e : events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); } events.clear(); }
    ~LaunchTimer() { clear(); }
};

#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
/* make the initialised barrier visible to the async (TMA) proxy */
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
/* 1-D bulk copy global -> shared through the TMA engine (SASS: UBLKCP); 16-byte aligned src/dst/size */
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
/* bounded wait: a lost transaction traps instead of hanging the GPU */
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 22)) __trap();
}
#endif

}  // namespace lbad
#endif
