// Shared helpers for the lwb_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/lwb_b200.h"

namespace lwb {

void set_error(const char* fmt, ...);

#define LWB_CHECK_ARG(cond, msg)                                                   \
    do { if (!(cond)) { lwb::set_error("%s: %s", __func__, msg); return LWB_E_INVALID; } } while (0)

#define LWB_CUDA_OK(expr)                                                          \
    do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) {                       \
        lwb::set_error("%s: %s -> %s", __func__, #expr, cudaGetErrorString(e__));  \
        return LWB_E_CUDA; } } while (0)

#define LWB_LAUNCH_OK()                                                            \
    do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) {           \
        lwb::set_error("%s: launch failed -> %s", __func__, cudaGetErrorString(e__)); \
        return LWB_E_CUDA; } } while (0)

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

int sm_count();

// x ~= hi + lo with hi = fp16(x), lo = fp16(x - hi): the 2-term operand split of the conv engine.
__device__ __forceinline__ void split_half(float x, __half& hi, __half& lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
}

}  // namespace lwb
