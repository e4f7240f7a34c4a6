"""HMR image encoder + theta regressor (networks/hmr.py:255-300) on the conv engine against
  * tests/golden/hmr.npz -- theta / encoder features the REFERENCE module produced on the same seeded weights and images
    (tests/golden/make_hmr_golden.py), and
  * the functional restatement oracle/hmr_ref.py run here on CPU.
Tolerance: 1e-3 max-abs on theta (the 85 SMPL parameters the rest of the path consumes)."""
import os

import numpy as np
import pytest
import torch

from impersonator_b200 import synthetic as S
from impersonator_b200.hmr import HumanModelRecovery
from oracle import hmr_ref

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "hmr.npz")


@pytest.fixture(scope="module")
def hmr(cuda):
    torch.set_grad_enabled(False)
    net = HumanModelRecovery(smpl_model=S.synthetic_smpl_model(seed=3)).eval()
    sd = S.synthetic_hmr_state(net.state_dict())
    full = dict(net.state_dict())
    full.update(sd)
    net.load_state_dict(full, strict=True)
    return net.to(cuda).eval(), sd


def test_state_dict_keys_match_reference(hmr):
    net, _ = hmr
    g = np.load(GOLD)
    mine = {k: str(tuple(v.shape)) for k, v in net.state_dict().items()}
    ref = dict(zip(g["keys"].tolist(), g["shapes"].tolist()))
    assert mine == ref


@pytest.mark.parametrize("mode", ["fp16f8", "fp16x3"])
def test_theta_matches_reference_golden_and_oracle(cuda, hmr, monkeypatch, mode):
    net, sd = hmr
    monkeypatch.setenv("LWB_PRECISION", mode)
    g = np.load(GOLD)
    x = S.synthetic_hmr_inputs(3)
    theta = net(x.to(cuda)).cpu()
    d_gold = np.abs(theta.numpy() - g["theta"]).max()
    ref = hmr_ref.forward(x, sd)
    d_or = (theta - ref).abs().max().item()
    print("%s: theta vs reference golden %.3e, vs oracle %.3e (theta range %.2f..%.2f)" % (mode, d_gold, d_or, ref.min(), ref.max()))
    assert d_gold < 1e-3 and d_or < 1e-3
    one = net(x[1:2].to(cuda)).cpu()                                   # batch 1 (how models/imitator.py:95,273 calls it)
    assert (one - theta[1:2]).abs().max().item() < 1e-4
    det = net.get_details(theta.to(cuda))
    assert det["verts"].shape == (3, 6890, 3) and det["j2d"].shape == (3, 19, 2)
