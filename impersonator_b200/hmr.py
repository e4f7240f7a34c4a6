"""``HumanModelRecovery`` (networks/hmr.py:255-330) -- the part the per-frame path needs.

``get_details(theta)`` (:302-330): theta [N,85] = cam(3) | pose(72) | shape(10) -> cam, pose, shape, verts, j3d, j2d through
the SMPL kernels.  The image encoder (pre-activation ResNet-50 + iterative regressor, :275-300) is SURVEY.md 8f rank 3
and is not built: ``forward`` fails loudly; pass ``tgt_smpls`` (as run_imitator.py does for pre-computed SMPL files)
or inject the reference's HMR for that step.
"""
import torch.nn as nn

from ._lib import LwbError
from .smpl import SMPL


class HumanModelRecovery(nn.Module):
    def __init__(self, smpl_pkl_path=None, feature_dim=2048, theta_dim=85, iterations=3, smpl_model=None):
        super(HumanModelRecovery, self).__init__()
        self.smpl = SMPL(pkl_path=smpl_pkl_path, model=smpl_model)
        self.feature_dim = feature_dim
        self.theta_dim = theta_dim
        self.iterations = iterations

    def forward(self, inputs):
        raise LwbError("the HMR image encoder is not part of this library (SURVEY.md 8f rank 3): "
                       "pass SMPL vectors (tgt_smpls / src_smpl) or inject the reference HMR")

    def get_details(self, theta):
        cam = theta[:, 0:3].contiguous()
        pose = theta[:, 3:75].contiguous()
        shape = theta[:, 75:].contiguous()
        verts, j3d, rs = self.smpl(beta=shape, theta=pose, get_skin=True, cam=cam)
        return {'theta': theta, 'cam': cam, 'pose': pose, 'shape': shape, 'verts': verts,
                'j2d': self.smpl.j2d, 'j3d': j3d}
