// vt_common.cuh -- shared helpers for libvtseg (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/vtseg.h"

namespace vt {

void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline int cuda_fail(cudaError_t e, const char *what) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return VT_ERR_CUDA;
}

#define VT_CUDA(call)                                         \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return vt::cuda_fail(e__, #call); \
    } while (0)

// Count a launch and surface launch-time errors (bad config, missing image) immediately.
#define VT_LAUNCHED(name)                                      \
    do {                                                       \
        vt::g_launches.fetch_add(1, std::memory_order_relaxed); \
        cudaError_t e__ = cudaGetLastError();                  \
        if (e__ != cudaSuccess) return vt::cuda_fail(e__, name); \
    } while (0)

#define VT_MAX_DEVICES 64
int current_device();   // ordinal of the calling thread's device (per-device caches index by it)
int sm_count();

// ---- device-side helpers ------------------------------------------------------------------------------
#ifdef __CUDACC__
// Streaming 128-bit load that does not allocate in L1 (each byte is consumed once).
__device__ __forceinline__ uint4 ld_stream_u4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_u4(void *p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// sum of |a_i - b_i| over the four bytes of a and b, added to acc
__device__ __forceinline__ uint32_t sad4(uint32_t a, uint32_t b, uint32_t acc) {
    uint32_t d = __vabsdiffu4(a, b);
    return __dp4a(d, 0x01010101u, acc);
}

// mbarrier + TMA (cp.async.bulk.tensor) wrappers
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 3-D tiled load: coordinates (x bytes, y row, z frame)
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const void *tmap, uint64_t *bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
        "[%2];" ::"r"(smem_u32(smem_dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}
#endif

}  // namespace vt
