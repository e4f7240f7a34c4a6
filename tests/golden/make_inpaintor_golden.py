"""Generates tests/golden/inpaintor.npz -- run ONLY in the build container (needs /root/reference).
Imports the reference's networks/inpaintor.py, loads deterministic weights
(impersonator_b200.synthetic.fill_state_dict seed 3, conv std 0.05), runs InpaintSANet.forward
(networks/inpaintor.py:178-202) on a synthetic image + mask, checks oracle/inpaintor_ref.py against
it on the full tensors and stores strided slices."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
for m in ("ipdb", "h5py"):
    sys.modules.setdefault(m, types.ModuleType(m))
sys.path.insert(0, "/root/reference")
from networks.inpaintor import InpaintSANet                     # noqa: E402  (the reference)
from impersonator_b200 import synthetic as S                    # noqa: E402
from oracle import inpaintor_ref as R                           # noqa: E402


def inputs():
    img = S.synthetic_source(256, seed=5)
    ys, xs = torch.meshgrid(torch.linspace(-1, 1, 256), torch.linspace(-1, 1, 256), indexing="ij")
    mask = (((xs / 0.4) ** 2 + (ys / 0.8) ** 2) < 1).float()[None, None]
    return img, mask


def main():
    torch.set_grad_enabled(False)
    net = InpaintSANet(c_dim=4).eval()
    sd = S.fill_state_dict(net.state_dict(), seed=3, conv_std=0.05)
    net.load_state_dict(sd)
    img, mask = inputs()
    ref = net(img, mask)
    mine = R.forward(img, mask, sd)
    out = {"keys": np.array(sorted(sd.keys()))}
    for name, a, b in zip(("coarse", "x", "comp"), ref, mine):
        d = (a - b).abs().max().item()
        print("%-6s restatement-vs-reference max-abs %.3g  (|ref| max %.3g)" % (name, d, a.abs().max().item()))
        assert d < 1e-5
        out[name] = a[:, :, 3::8, 5::8].contiguous().numpy()
    np.savez_compressed(os.path.join(HERE, "inpaintor.npz"), **out)


if __name__ == "__main__":
    main()
