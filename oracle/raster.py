"""TEST INFRASTRUCTURE ONLY -- ctypes front-ends of the rasterizer oracles.

* ``forward_face_index_map_cpu``: oracle/raster_ref.c (C restatement of
  rasterize_cuda_kernel.cu:40-186), bit-level model of the reference kernels.
* ``forward_face_index_map_gpu_ref``: oracle/_ref/libnmr_ref.so = the reference's own
  kernels + launcher (rasterize_cuda_kernel.cu:613-668) compiled unmodified for sm_100a.
Both mirror rasterize.py:50-52,164-169: outputs pre-filled with -1 / 0 / far, kernel row order
(row 0 = bottom); ``rasterize_fim_wim`` adds the flips of rasterize.py:334-338.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_cpu = None
_gpu = None


def build():
    """Compile the C restatement (and the GPU oracle when /root/reference exists)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "liblwb_oracle.so"])
    subprocess.check_call(["bash", os.path.join(_HERE, "build_ref.sh")],
                          stdout=subprocess.DEVNULL)


def _cpu_lib():
    global _cpu
    if _cpu is None:
        path = os.path.join(_HERE, "liblwb_oracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-s", "-C", _HERE, "liblwb_oracle.so"])
        _cpu = ctypes.CDLL(path)
        _cpu.lwb_oracle_forward_face_index_map.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _cpu.lwb_oracle_forward_face_index_map.restype = None
        _cpu.lwb_oracle_num_threads.restype = ctypes.c_int
    return _cpu


def num_threads():
    return int(_cpu_lib().lwb_oracle_num_threads())


def forward_face_index_map_cpu(faces, image_size, near=0.1, far=100.0):
    """faces f32[B,F,3,3] (numpy) -> fim i32[B,H,W], wim f32[B,H,W,3], depth f32[B,H,W],
    faces_inv f32[B,F,3,3]; kernel row order (no flip)."""
    faces = np.ascontiguousarray(faces, dtype=np.float32)
    B, F = faces.shape[:2]
    s = int(image_size)
    fim = np.full((B, s, s), -1, np.int32)
    wim = np.zeros((B, s, s, 3), np.float32)
    depth = np.full((B, s, s), far, np.float32)
    finv = np.zeros((B, F, 3, 3), np.float32)
    _cpu_lib().lwb_oracle_forward_face_index_map(
        faces.ctypes.data, B, F, s, near, far,
        fim.ctypes.data, wim.ctypes.data, depth.ctypes.data, finv.ctypes.data)
    return fim, wim, depth, finv


def rasterize_fim_wim(faces, image_size, near=0.1, far=100.0):
    """nr.rasterize_face_index_map_and_weight_map(faces, image_size, False) (rasterize.py:543-571):
    the kernel outputs flipped along H."""
    fim, wim, depth, _ = forward_face_index_map_cpu(faces, image_size, near, far)
    return fim[:, ::-1].copy(), wim[:, ::-1].copy(), depth[:, ::-1].copy()


def gpu_ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libnmr_ref.so"))


def forward_face_index_map_gpu_ref(faces_t, image_size, near=0.1, far=100.0):
    """Run the reference's own CUDA kernels (GPU oracle).  faces_t: torch cuda f32[B,F,3,3].
    Returns torch cuda tensors (fim, wim, depth, faces_inv), kernel row order."""
    import torch
    global _gpu
    if _gpu is None:
        _gpu = ctypes.CDLL(os.path.join(_HERE, "_ref", "libnmr_ref.so"))
        _gpu.nmr_ref_forward_face_index_map.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _gpu.nmr_ref_forward_face_index_map.restype = ctypes.c_int
    faces_t = faces_t.contiguous().float()
    B, F = faces_t.shape[:2]
    s = int(image_size)
    dev = faces_t.device
    fim = torch.full((B, s, s), -1, dtype=torch.int32, device=dev)
    wim = torch.zeros((B, s, s, 3), dtype=torch.float32, device=dev)
    depth = torch.full((B, s, s), far, dtype=torch.float32, device=dev)
    finv = torch.zeros((B, F, 3, 3), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    err = _gpu.nmr_ref_forward_face_index_map(
        faces_t.data_ptr(), B, F, s, near, far,
        fim.data_ptr(), wim.data_ptr(), depth.data_ptr(), finv.data_ptr())
    if err != 0:
        raise RuntimeError("reference rasterizer failed: cuda error %d" % err)
    return fim, wim, depth, finv
