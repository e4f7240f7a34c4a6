"""Drop-in for the pybind11 module ``neural_renderer.cuda.rasterize``
(thirdparty/neural_renderer/neural_renderer/cuda/rasterize_cuda.cpp:194-200).

Only ``forward_face_index_map`` is on the inference path; it keeps the reference contract
(rasterize_cuda.cpp:70-95): the caller allocates and pre-fills the maps, the kernel writes covered
pixels in place (native row order, row 0 = bottom) and the same tensors are returned.
"""
from ... import kernels as K
from ..._lib import LwbError


def _check(t, name):
    # rasterize_cuda.cpp:66-68 CHECK_INPUT
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)


def forward_face_index_map(faces, face_index_map, weight_map, depth_map, face_inv_map, faces_inv,
                           image_size, near, far, return_rgb, return_alpha, return_depth):
    for t, n in ((faces, "faces"), (face_index_map, "face_index_map"), (weight_map, "weight_map"),
                 (depth_map, "depth_map"), (face_inv_map, "face_inv_map"), (faces_inv, "faces_inv")):
        _check(t, n)
    if return_depth:
        raise LwbError("face_inv_map (return_depth) is only consumed by the backward pass: out of scope")
    K.raster_forward_face_index_map(faces, face_index_map, weight_map, depth_map, int(image_size),
                                    near=float(near), far=float(far),
                                    faces_inv=faces_inv if faces_inv.numel() == faces.numel() else None)
    return [face_index_map, weight_map, depth_map, face_inv_map]


def _out_of_scope(*a, **k):
    raise LwbError("texture sampling / backward kernels are outside the inference hot path (SURVEY.md section 8)")


forward_texture_sampling = _out_of_scope
backward_pixel_map = _out_of_scope
backward_textures = _out_of_scope
backward_depth_map = _out_of_scope
