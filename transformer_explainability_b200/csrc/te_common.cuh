// Shared device helpers for the sm_100a transformer-attribution kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define TE_OK 0
#define TE_ERR_ARG (-1)
#define TE_ERR_WORKSPACE (-2)
#define TE_ERR_CUDA (-3)
#define TE_ERR_UNSUPPORTED (-4)

#define TE_CUDA_CHECK_LAUNCH()                                   \
    do {                                                         \
        te_count_launch();                                       \
        cudaError_t e__ = cudaGetLastError();                    \
        if (e__ != cudaSuccess) { te_set_last_error(cudaGetErrorString(e__)); return TE_ERR_CUDA; } \
    } while (0)

#define TE_TRY(x)                                                \
    do { int r__ = (x); if (r__ != TE_OK) return r__; } while (0)

void te_set_last_error(const char* msg);
void te_count_launch();

// safe_divide of the reference (modules/layers_ours.py:10-13):
//   den = clamp(b,min=eps) + clamp(b,max=eps)  ( == b + eps ) ; den += eps where den == 0 ;
//   out = a / den * (b != 0)
template <typename T>
__device__ __forceinline__ T te_sd(T a, T b) {
    const T eps = (T)1e-9;
    T den = b + eps;
    den = (den == (T)0) ? eps : den;
    return (a / den) * ((b != (T)0) ? (T)1 : (T)0);
}

__device__ __forceinline__ float te_gelu(float x) {           // exact (erf) GELU, nn.GELU default
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float te_gelu_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
    return cdf + x * pdf;
}

__device__ __forceinline__ float te_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double te_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float te_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

static inline int te_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
