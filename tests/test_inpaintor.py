"""InpaintSANet mirror (once-per-source background network, networks/inpaintor.py)."""
import os
import sys

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLD)


def _inputs():
    from impersonator_b200 import synthetic as S
    img = S.synthetic_source(256, seed=5)
    ys, xs = torch.meshgrid(torch.linspace(-1, 1, 256), torch.linspace(-1, 1, 256), indexing="ij")
    mask = (((xs / 0.4) ** 2 + (ys / 0.8) ** 2) < 1).float()[None, None]
    return img, mask


def test_restatement_and_keys_match_reference_golden():
    from impersonator_b200 import synthetic as S
    from impersonator_b200.inpaintor import InpaintSANet
    from oracle import inpaintor_ref as R
    torch.set_grad_enabled(False)
    g = np.load(os.path.join(GOLD, "inpaintor.npz"))
    net = InpaintSANet(c_dim=4)
    assert sorted(net.state_dict().keys()) == list(g["keys"])                 # 322 reference keys
    sd = S.fill_state_dict(net.state_dict(), seed=3, conv_std=0.05)
    img, mask = _inputs()
    coarse, x, comp = R.forward(img, mask, sd)
    for name, t in (("coarse", coarse), ("x", x), ("comp", comp)):
        assert np.abs(t[:, :, 3::8, 5::8].numpy() - g[name]).max() < 1e-5


@pytest.mark.gpu
def test_inpaintor_gpu_matches_reference_golden(cuda):
    from impersonator_b200 import synthetic as S
    from impersonator_b200.inpaintor import InpaintSANet
    torch.set_grad_enabled(False)
    g = np.load(os.path.join(GOLD, "inpaintor.npz"))
    net = InpaintSANet(c_dim=4)
    net.load_state_dict(S.fill_state_dict(net.state_dict(), seed=3, conv_std=0.05))
    net = net.to(cuda).eval()
    img, mask = _inputs()
    coarse, x, comp = net(img.to(cuda), mask.to(cuda))
    for name, t in (("coarse", coarse), ("x", x), ("comp", comp)):
        d = np.abs(t[:, :, 3::8, 5::8].cpu().numpy() - g[name]).max()
        print("inpaintor %s vs reference golden: %.3e" % (name, d))
        assert d < 1e-3
    only_x = net(img.to(cuda), mask.to(cuda), only_x=True)                    # models/imitator.py:125
    assert torch.equal(only_x, x)
