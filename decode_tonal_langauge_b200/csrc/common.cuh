// Shared helpers for libecog_sm100 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/ecog_sm100.h"

namespace ecog {

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs

extern thread_local char g_err[512];
extern thread_local int64_t g_launches;
extern thread_local char g_launch_log[4096];     // names of the most recent launches, comma separated
extern thread_local int g_launch_log_len;

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

inline int check_launch(const char* what) {
    ++g_launches;
    {   // bookkeeping for tests / bench: which kernel families ran (ecog_launch_log)
        const int n = (int)strlen(what);
        if (g_launch_log_len + n + 2 < (int)sizeof(g_launch_log)) {
            if (g_launch_log_len) g_launch_log[g_launch_log_len++] = ',';
            memcpy(g_launch_log + g_launch_log_len, what, n);
            g_launch_log_len += n;
            g_launch_log[g_launch_log_len] = 0;
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ECOG_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return ECOG_OK;
}

#define ECOG_TRY(expr)                      \
    do {                                    \
        int _rc = (expr);                   \
        if (_rc != ECOG_OK) return _rc;     \
    } while (0)

#define ECOG_CUDA(expr)                                                                   \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess)                                                            \
            return ::ecog::fail(ECOG_E_CUDA, "%s: %s", #expr, cudaGetErrorString(_e));    \
    } while (0)

// Opt a kernel instance in to `smem` bytes of dynamic shared memory: issued once per kernel, device
// and size (grow only), not on every launch.
template <auto K>
inline int smem_attr(size_t smem) {
    static size_t cur[64] = {};
    int dev = 0;
    ECOG_CUDA(cudaGetDevice(&dev));
    const bool track = dev >= 0 && dev < 64;
    if (!track || smem > cur[dev]) {
        ECOG_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (track) cur[dev] = smem;
    }
    return ECOG_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;   // read-once stream: do not allocate in L1
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8_zfill(void* smem, const void* gmem, bool valid) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    int n = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(s), "l"(gmem), "r"(n) : "memory");
}
// 4-byte copy with zero fill when !valid (src-size 0)
__device__ __forceinline__ void cp_async4_zfill(void* smem, const void* gmem, bool valid) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    int n = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(s), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, bool valid) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(s), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace ecog
