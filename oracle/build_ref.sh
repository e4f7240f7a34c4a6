#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  Compile the reference rasterizer (GPU oracle) from /root/reference.
# Runs only where /root/reference exists (the build container); the .so travels to the GPU box.
set -e
cd "$(dirname "$0")"
REF=${REF_ROOT:-/root/reference}
SRC=$REF/thirdparty/neural_renderer/neural_renderer/cuda/rasterize_cuda_kernel.cu
[ -f "$SRC" ] || { echo "reference not present; keeping prebuilt oracle/_ref"; exit 0; }
mkdir -p _ref
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC \
     -Ishim -DREF_KERNEL_CU="\"$SRC\"" ref_wrap.cu -o _ref/libnmr_ref.so
echo "built oracle/_ref/libnmr_ref.so"
