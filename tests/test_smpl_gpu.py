"""lwb_smpl_forward (csrc/smpl.cu) through the SMPL / HumanModelRecovery mirrors vs
  * outputs of the reference class networks.batch_smpl.SMPL (tests/golden/smpl.npz), and
  * the torch restatement oracle/smpl_ref.py on full tensors and other batch sizes.
fp32 throughout; tolerance 1e-5 absolute on metre-scale vertices (summation order differs from oneDNN's)."""
import os

import numpy as np
import pytest
import torch

from impersonator_b200 import synthetic as S
from impersonator_b200.hmr import HumanModelRecovery
from impersonator_b200.smpl import SMPL
from oracle import smpl_ref

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "smpl.npz")
TOL = 1e-5


def golden_theta():
    theta = S.synthetic_smpl_params(5, seed=17)
    theta[1, 3:75] = 0
    return theta


@pytest.mark.parametrize("rotate", [False, True])
def test_matches_reference_class_golden(cuda, rotate):
    torch.set_grad_enabled(False)
    g = np.load(GOLD)
    tag = "rot" if rotate else "std"
    smpl = SMPL(model=S.synthetic_smpl_model(seed=3), rotate=rotate).to(cuda)
    theta = golden_theta().to(cuda)
    verts, joints, Rs = smpl(beta=theta[:, 75:].contiguous(), theta=theta[:, 3:75].contiguous(), get_skin=True,
                             cam=theta[:, :3].contiguous())
    torch.cuda.synchronize()
    d = {"verts": np.abs(verts[:, ::13].cpu().numpy() - g["verts_" + tag]).max(),
         "joints": np.abs(joints.cpu().numpy() - g["joints_" + tag]).max(),
         "Rs": np.abs(Rs.cpu().numpy() - g["Rs_" + tag]).max(),
         "j2d": np.abs(smpl.j2d.cpu().numpy() - g["j2d_" + tag]).max()}
    print("smpl vs reference golden (%s): %s" % (tag, d))
    assert max(d.values()) < TOL
    only_joints = smpl(beta=theta[:, 75:].contiguous(), theta=theta[:, 3:75].contiguous())
    assert torch.equal(only_joints, joints)


@pytest.mark.parametrize("batch", [1, 8, 16, 19])
def test_matches_oracle_full_tensors(cuda, batch):
    torch.set_grad_enabled(False)
    dd = S.synthetic_smpl_model(seed=3)
    m = smpl_ref.model_tensors(dd)
    hmr = HumanModelRecovery(smpl_model=dd).to(cuda)
    theta = S.synthetic_smpl_params(batch, seed=100 + batch)
    ref = smpl_ref.get_details(m, theta)
    _, _, _, Jt_ref = smpl_ref.forward(m, theta[:, 75:].contiguous(), theta[:, 3:75].contiguous())
    out = hmr.get_details(theta.to(cuda))
    torch.cuda.synchronize()
    for k in ("verts", "j3d", "j2d", "cam", "pose", "shape"):
        d = (out[k].cpu() - ref[k]).abs().max().item()
        assert d < TOL, (k, d)
    assert (hmr.smpl.J_transformed.cpu() - Jt_ref).abs().max().item() < TOL


def test_errors_are_loud(cuda):
    from impersonator_b200._lib import LwbError
    hmr = HumanModelRecovery(smpl_model=S.synthetic_smpl_model(seed=3)).to(cuda)
    with pytest.raises(LwbError):
        hmr(torch.zeros(1, 3, 224, 224, device=cuda))
    with pytest.raises(LwbError):
        hmr.smpl(beta=torch.zeros(2, 9, device=cuda), theta=torch.zeros(2, 72, device=cuda))
    with pytest.raises(LwbError):
        hmr.smpl(beta=torch.zeros(2, 10), theta=torch.zeros(2, 72))
