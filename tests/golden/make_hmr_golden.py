"""Generates tests/golden/hmr.npz -- run ONLY in the build container (needs /root/reference).

Imports the REFERENCE ``networks.hmr.HumanModelRecovery`` (h5py / ipdb stubbed; SMPL from a synthetic smpl_model.pkl, as
tests/golden/make_smpl_golden.py does), loads impersonator_b200.synthetic.fill_state_dict(seed=4, conv_std='he') into its
resnet + regressor (the weights are regenerated from the seed wherever they are needed: 27 M parameters do not go into
git), runs ``forward`` on seeded images and stores theta + the 2048 encoder features.  Also checks the functional
restatement oracle/hmr_ref.py against the reference module on the full outputs."""
import os
import pickle
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
for stub in ("ipdb", "h5py"):
    sys.modules.setdefault(stub, types.ModuleType(stub))
sys.path.insert(0, "/root/reference")

from impersonator_b200 import synthetic as S          # noqa: E402
from networks.hmr import HumanModelRecovery           # noqa: E402  (the reference)
from oracle import hmr_ref                            # noqa: E402


def main():
    torch.set_grad_enabled(False)
    with tempfile.NamedTemporaryFile(suffix=".pkl", delete=False) as fp:
        pickle.dump(S.synthetic_smpl_model(seed=3), fp, protocol=2)
    net = HumanModelRecovery(fp.name).eval()
    os.unlink(fp.name)
    sd = S.synthetic_hmr_state(net.state_dict())
    full = dict(net.state_dict())
    full.update(sd)
    net.load_state_dict(full, strict=True)
    x = S.synthetic_hmr_inputs(3)
    theta = net(x)
    feat = hmr_ref.encoder(x, sd)
    mine = hmr_ref.forward(x, sd)
    d = (theta - mine).abs().max().item()
    print("restatement-vs-reference theta max-abs %.3g; theta range [%.3f, %.3f]; feature mean %.3f max %.3f"
          % (d, theta.min(), theta.max(), feat.mean(), feat.max()))
    assert d < 1e-5
    keys = sorted(k for k in net.state_dict().keys())
    np.savez_compressed(os.path.join(HERE, "hmr.npz"), theta=theta.numpy(), features=feat.numpy(),
                        keys=np.array(keys), shapes=np.array([str(tuple(net.state_dict()[k].shape)) for k in keys]))
    print("wrote hmr.npz", theta.shape, feat.shape, len(keys), "keys")


if __name__ == "__main__":
    main()
