"""Generates tests/golden/smpl.npz by running the REFERENCE class networks.batch_smpl.SMPL (imported from
/root/reference, CPU) on the synthetic model file (impersonator_b200.synthetic.synthetic_smpl_model, written as a
protocol-2 pickle exactly like smpl_model.pkl) and seeded parameters.  Run in the build container only."""
import os
import pickle
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
for stub in ("ipdb", "h5py"):
    sys.modules.setdefault(stub, types.ModuleType(stub))
sys.path.insert(0, "/root/reference")

from impersonator_b200 import synthetic as S  # noqa: E402
from networks.batch_smpl import SMPL, batch_orth_proj_idrot  # noqa: E402

torch.set_grad_enabled(False)
dd = S.synthetic_smpl_model(seed=3)
with tempfile.NamedTemporaryFile(suffix=".pkl", delete=False) as fp:
    pickle.dump(dd, fp, protocol=2)
out = {}
for rotate in (False, True):
    smpl = SMPL(fp.name, rotate=rotate)
    theta = S.synthetic_smpl_params(5, seed=17)
    theta[1, 3:75] = 0                                   # rest pose: Rodrigues at angle ~ 1e-8
    cam, pose, shape = theta[:, :3], theta[:, 3:75].contiguous(), theta[:, 75:].contiguous()
    verts, joints, Rs = smpl(beta=shape, theta=pose, get_skin=True)
    tag = "rot" if rotate else "std"
    out["verts_" + tag] = verts[:, ::13].numpy()         # every 13th vertex (530 of 6890) keeps the file small
    out["joints_" + tag] = joints.numpy()
    out["Rs_" + tag] = Rs.numpy()
    out["j2d_" + tag] = batch_orth_proj_idrot(joints, cam).numpy()
os.unlink(fp.name)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "smpl.npz"), **out)
print({k: v.shape for k, v in out.items()})
