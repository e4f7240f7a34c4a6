// TEST INFRASTRUCTURE ONLY (oracle/): a minimal stand-in for <ATen/ATen.h> so the
// reference file thirdparty/neural_renderer/neural_renderer/cuda/rasterize_cuda_kernel.cu
// compiles UNMODIFIED, from where it lies under /root/reference, with plain nvcc
// (its torch-1.2 era API -- Tensor::data<T>(), x.type(), AT_DISPATCH_FLOATING_TYPES --
// no longer exists in torch 2.11).  Only what that one file touches is provided.
#pragma once
#include <cstdint>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

namespace at {
struct Tensor {
    void*   ptr = nullptr;
    int64_t sizes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int     nd = 0;
    Tensor() {}
    Tensor(void* p, std::initializer_list<int64_t> s) : ptr(p), nd(0) {
        for (auto v : s) sizes[nd++] = v;
    }
    int64_t size(int i) const { return sizes[i]; }
    int64_t numel() const { int64_t n = 1; for (int i = 0; i < nd; ++i) n *= sizes[i]; return n; }
    int     type() const { return 0; }
    template <typename T> T* data() const { return reinterpret_cast<T*>(ptr); }
};
}  // namespace at

// float32 is the only dtype the path uses (rasterize.py:50-52 allocates FloatTensor).
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...) \
    { (void)(TYPE); using scalar_t = float; __VA_ARGS__(); }
