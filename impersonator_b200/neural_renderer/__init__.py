"""Host-side mirror of the slice of ``neural_renderer`` the hot path uses
(thirdparty/neural_renderer/neural_renderer/__init__.py): ``look_at``, ``vertices_to_faces``,
``rasterize_face_index_map_and_weight_map`` (+ ``rasterize_face_index_map``,
``rasterize_silhouettes``, ``rasterize_depth``), and the native entry point
``neural_renderer.cuda.rasterize.forward_face_index_map``.

Texture sampling, lighting, obj I/O and every backward kernel are out of scope (SURVEY.md 8:
inference never requests rgb/alpha and runs under no_grad) and raise if called.
"""
from .look_at import look_at
from .vertices_to_faces import vertices_to_faces
from .rasterize import (rasterize_face_index_map_and_weight_map, rasterize_face_index_map,
                        rasterize_silhouettes, rasterize_depth, rasterize_rgbad,
                        DEFAULT_IMAGE_SIZE, DEFAULT_ANTI_ALIASING, DEFAULT_NEAR, DEFAULT_FAR, DEFAULT_EPS)
from . import cuda

__version__ = '1.1.3-lwb_b200'
