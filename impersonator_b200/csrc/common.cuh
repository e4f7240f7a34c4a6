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

int sm_count();                      // of the current device
constexpr int kMaxDevices = 64;
int device_slot();                   // current device ordinal (clamped to kMaxDevices - 1): index of per-device caches

// Programmatic dependent launch (LWB_PDL, default on): kernels launched through launch_pdl may be scheduled while the
// previous kernel of the stream drains; each of them executes pdl_wait() before its first global-memory access and
// pdl_trigger() to let its own successor do the same.  Both are no-ops for a kernel launched the plain way.
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): one full 32-byte sector per lane and instruction instead of two
// 16-byte halves -- half the L1/L2 transactions of the fp32 activation streams.  Addresses must be 32-byte aligned.
__device__ __forceinline__ void ldg_f32x8(const float* p, float* v) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
__device__ __forceinline__ void stg_f32x8(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}

// x ~= hi + lo with hi = fp16(x), lo = fp16(x - hi): the 2-term operand split of the conv engine.
__device__ __forceinline__ void split_half(float x, __half& hi, __half& lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
}

}  // namespace lwb
