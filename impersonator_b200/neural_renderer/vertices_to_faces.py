"""``nr.vertices_to_faces``: gather the three vertices of every face.
Mirrors thirdparty/neural_renderer/neural_renderer/vertices_to_faces.py:4-22 (same result).
On the hot path this gather is fused into the raster kernel's vertex fetch (lwb_correspond)."""
import torch


def vertices_to_faces(vertices, faces):
    """vertices f32[B,V,3], faces int[B,F,3] -> f32[B,F,3,3]."""
    if vertices.dim() != 3 or faces.dim() != 3 or vertices.shape[-1] != 3 or faces.shape[-1] != 3:
        raise ValueError("expected vertices [B,V,3] and faces [B,F,3]")
    if vertices.shape[0] != faces.shape[0]:
        raise ValueError("batch sizes differ")
    B, V, _ = vertices.shape
    base = torch.arange(B, device=vertices.device, dtype=torch.long).view(B, 1, 1) * V
    return vertices.reshape(B * V, 3)[faces.long() + base]
