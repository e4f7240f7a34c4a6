// TEST INFRASTRUCTURE ONLY (oracle/).  Builds the GPU oracle: the reference's own
// rasterizer kernels + its own launcher (rasterize_cuda_kernel.cu:613-668: 512 threads,
// legacy default stream), compiled from /root/reference by the same nvcc with default
// flags (-fmad=true), behind a C ABI.  Output: oracle/_ref/libnmr_ref.so (git-ignored).
// No reference source is copied: the file is #included from where it lies (-DREF_KERNEL_CU).
#include <ATen/ATen.h>          // oracle/shim
#include REF_KERNEL_CU

extern "C" int nmr_ref_forward_face_index_map(
        const float* faces, int batch, int num_faces, int image_size, float near, float far,
        int32_t* face_index_map, float* weight_map, float* depth_map, float* faces_inv) {
    at::Tensor t_faces((void*)faces, {batch, num_faces, 3, 3});
    at::Tensor t_fim(face_index_map, {batch, image_size, image_size});
    at::Tensor t_wim(weight_map, {batch, image_size, image_size, 3});
    at::Tensor t_depth(depth_map, {batch, image_size, image_size});
    float* dummy = nullptr;
    cudaMalloc(&dummy, sizeof(float));
    at::Tensor t_fimap(dummy, {1});
    at::Tensor t_finv(faces_inv, {batch, num_faces, 3, 3});
    forward_face_index_map_cuda(t_faces, t_fim, t_wim, t_depth, t_fimap, t_finv,
                                image_size, near, far, 0, 0, 0);
    cudaError_t e = cudaDeviceSynchronize();
    cudaFree(dummy);
    return (int)e;
}
