// Shared helpers for the pixelssl_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/pixelssl_b200.h"

#define PXL_NUM_SMS 148

extern "C" void pxl_count_launch_(int n);

#define PXL_CHECK_LAUNCH()                                   \
    do {                                                     \
        pxl_count_launch_(1);                                \
        cudaError_t e__ = cudaPeekAtLastError();             \
        if (e__ != cudaSuccess) return (int)e__;             \
    } while (0)

static inline int64_t pxl_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// streaming (read-once / write-once) 128-bit accesses that do not pollute L1
__device__ __forceinline__ float4 ld_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// power-of-two scale s = 2^(target_log2 - e) with amax <= 2^e: amax * s lands in (2^(target-1), 2^target]
// (fp16 pairs of gradient tensors, csrc/h16_prep.cu); 1 for a zero / non-finite absmax
__device__ __forceinline__ float pxl_pow2_scale(float amax, int target_log2) {
    if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
    int e;
    frexpf(amax, &e);
    int k = target_log2 - e;
    if (k > 126) k = 126;
    if (k < -126) k = -126;
    return ldexpf(1.f, k);
}
