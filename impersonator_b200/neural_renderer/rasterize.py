"""Mirror of thirdparty/neural_renderer/neural_renderer/rasterize.py for the forward,
texture-less products the inference path consumes: face-index map, weight map, silhouette, depth.

``RasterizeFunction.forward`` (rasterize.py:22-98) allocates + pre-fills the maps and calls the
native kernel; ``rasterize_rgbad`` (rasterize.py:257-358) adds the vertical flips and the optional
2x super-sampling.  Here the flip is folded into the kernel's store (flip_rows=1).
"""
import torch
import torch.nn.functional as F

from .. import kernels as K
from .._lib import LwbError

DEFAULT_IMAGE_SIZE = 256
DEFAULT_ANTI_ALIASING = True
DEFAULT_NEAR = 0.1
DEFAULT_FAR = 100
DEFAULT_EPS = 1e-4
DEFAULT_BACKGROUND_COLOR = (0, 0, 0)


def _raster(faces, image_size, near, far):
    if not faces.is_cuda:
        raise TypeError('Rasterize module supports only cuda Tensors')       # rasterize.py:250-251
    faces = faces.detach().float().contiguous()
    B = faces.shape[0]
    dev = faces.device
    fim = torch.full((B, image_size, image_size), -1, dtype=torch.int32, device=dev)     # rasterize.py:50
    wim = torch.zeros((B, image_size, image_size, 3), dtype=torch.float32, device=dev)   # :51
    depth = torch.full((B, image_size, image_size), float(far), dtype=torch.float32, device=dev)   # :52
    K.raster_forward_face_index_map(faces, fim, wim, depth, image_size, near=float(near), far=float(far), flip_rows=True)
    return fim, wim, depth


def rasterize_rgbad(faces, textures=None, image_size=DEFAULT_IMAGE_SIZE, anti_aliasing=DEFAULT_ANTI_ALIASING,
                    near=DEFAULT_NEAR, far=DEFAULT_FAR, eps=DEFAULT_EPS, background_color=DEFAULT_BACKGROUND_COLOR,
                    return_rgb=True, return_alpha=True, return_depth=True, return_fim=True, return_weight=True):
    if return_rgb or textures is not None:
        raise LwbError("textured rendering (return_rgb) is outside the inference hot path (SURVEY.md section 8)")
    size = image_size * 2 if anti_aliasing else image_size
    fim, wim, depth = _raster(faces, size, near, far)
    alpha = (fim >= 0).float() if return_alpha else None                      # rasterize.py:191 forward_alpha_map
    if anti_aliasing:                                                         # rasterize.py:340-347
        if return_alpha:
            alpha = F.avg_pool2d(alpha[:, None, :, :], kernel_size=(2, 2))[:, 0]
        if return_depth:
            depth = F.avg_pool2d(depth[:, None, :, :], kernel_size=(2, 2))[:, 0]
    return {'rgb': None, 'alpha': alpha, 'depth': depth if return_depth else None,
            'face_index_map': fim if return_fim else None, 'weight_map': wim if return_fim else None}


def rasterize_face_index_map_and_weight_map(faces, image_size=DEFAULT_IMAGE_SIZE, anti_aliasing=DEFAULT_ANTI_ALIASING,
                                            near=DEFAULT_NEAR, far=DEFAULT_FAR, eps=DEFAULT_EPS):
    """rasterize.py:543-571 -> (face_index_map i32[B,H,W], weight_map f32[B,H,W,3]), top row first."""
    size = image_size * 2 if anti_aliasing else image_size
    fim, wim, _ = _raster(faces, size, near, far)
    return fim, wim


def rasterize_face_index_map(faces, image_size=DEFAULT_IMAGE_SIZE, anti_aliasing=DEFAULT_ANTI_ALIASING,
                             near=DEFAULT_NEAR, far=DEFAULT_FAR, eps=DEFAULT_EPS):
    return rasterize_face_index_map_and_weight_map(faces, image_size, anti_aliasing, near, far, eps)[0]


def rasterize_silhouettes(faces, image_size=DEFAULT_IMAGE_SIZE, anti_aliasing=DEFAULT_ANTI_ALIASING,
                          near=DEFAULT_NEAR, far=DEFAULT_FAR, eps=DEFAULT_EPS):
    return rasterize_rgbad(faces, None, image_size, anti_aliasing, near, far, eps, None,
                           False, True, False, False, False)['alpha']


def rasterize_depth(faces, image_size=DEFAULT_IMAGE_SIZE, anti_aliasing=DEFAULT_ANTI_ALIASING,
                    near=DEFAULT_NEAR, far=DEFAULT_FAR, eps=DEFAULT_EPS):
    return rasterize_rgbad(faces, None, image_size, anti_aliasing, near, far, eps, None,
                           False, False, True, False, False)['depth']
