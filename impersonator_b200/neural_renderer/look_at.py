"""``nr.look_at`` (thirdparty/neural_renderer/neural_renderer/look_at.py:6-62), batched-safe.

The reference builds the camera axes with ``torch.cross`` without ``dim``, which picks the wrong
axis when the batch size is 3 on current torch; ``dim=-1`` is what the
author meant.  On the hot path the eye is the constant of utils/nmr.py:177, the rotation is the
identity, and the fused correspondence kernel applies ``v - eye`` itself.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _as_tensor(v, device):
    if isinstance(v, (list, tuple)):
        return torch.tensor(v, dtype=torch.float32, device=device)
    if isinstance(v, np.ndarray):
        return torch.from_numpy(v).to(device)
    return v.to(device)


def look_at(vertices, eye, at=[0, 0, 0], up=[0, 1, 0]):
    if vertices.ndimension() != 3:
        raise ValueError('vertices Tensor should have 3 dimensions')
    device = vertices.device
    at, up, eye = _as_tensor(at, device), _as_tensor(up, device), _as_tensor(eye, device)
    batch_size = vertices.shape[0]
    if eye.ndimension() == 1:
        eye = eye[None, :].repeat(batch_size, 1)
    if at.ndimension() == 1:
        at = at[None, :].repeat(batch_size, 1)
    if up.ndimension() == 1:
        up = up[None, :].repeat(batch_size, 1)
    z_axis = F.normalize(at - eye, eps=1e-5)
    x_axis = F.normalize(torch.linalg.cross(up, z_axis, dim=-1), eps=1e-5)
    y_axis = F.normalize(torch.linalg.cross(z_axis, x_axis, dim=-1), eps=1e-5)
    r = torch.cat((x_axis[:, None, :], y_axis[:, None, :], z_axis[:, None, :]), dim=1)
    if vertices.shape != eye.shape:
        eye = eye[:, None, :]
    vertices = vertices - eye
    return torch.matmul(vertices, r.transpose(1, 2))
