"""The CUDA correspondence path and the SMPLRenderer mirror against tests/golden/nmr.npz, i.e. against the
outputs of the REFERENCE's own utils/nmr.py methods (tests/golden/make_nmr_golden.py): fim bit-exact,
wim / cond / T / warped source image within 1e-5."""
import os

import numpy as np
import pytest
import torch

from impersonator_b200 import kernels as K
from impersonator_b200 import synthetic as S
from impersonator_b200.nmr import SMPLRenderer

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "nmr.npz")
TOL = 1e-5


def test_correspondence_matches_reference_nmr_methods(cuda):
    torch.set_grad_enabled(False)
    g = np.load(GOLD)
    v, f = S.uv_sphere()
    cam, verts = S.synthetic_frames(3, seed=41, base_verts=v)
    assert np.array_equal(cam.numpy(), g["cam"])
    tabs = S.synthetic_tables()
    src_img = S.synthetic_source(64).to(cuda)
    r = SMPLRenderer(image_size=64, faces=f.numpy(), map_fn=tabs["map_fn"], has_front=True,
                     front_map_fn=tabs["front_map_fn"], back_map_fn=tabs["back_map_fn"]).to(cuda)
    cam, verts = cam.to(cuda), verts.to(cuda)

    # source side: render_fim_wim + encode_fim + p2verts + get_vis_f2pts (models/imitator.py:98-110)
    s_f2v, s_fim, s_wim = r.render_fim_wim(cam[:1], verts[:1])
    assert np.array_equal(s_fim.cpu().numpy(), g["src_fim"])
    assert np.abs(s_wim.cpu().numpy() - g["src_wim"]).max() < TOL
    assert np.array_equal(s_f2v[:, ::7].cpu().numpy(), g["src_f2verts_sub"])          # projection is exact fp32
    s_cond, _ = r.encode_fim(cam[:1], verts[:1], fim=s_fim, transpose=True)
    assert np.array_equal(s_cond.cpu().numpy(), g["src_cond"])
    p2v = s_f2v[:, :, :, 0:2].clone()
    p2v[:, :, :, 1] *= -1
    vis = r.get_vis_f2pts(p2v, s_fim)
    assert np.array_equal(vis[:, ::7].cpu().numpy(), g["vis_sub"])
    assert np.array_equal((vis[0, :, 0, 0] != -2).nonzero()[:, 0].cpu().numpy(), g["vis_ids"])

    # target side: the fused kernel (render_fim_wim + encode_fim + cal_bc_transform + grid_sample + cat)
    for ac in (True, False):
        out = r.correspond(cam[1:], verts[1:], p2v.contiguous(), src_img, align_corners=ac, want_f2verts=True)
        assert np.array_equal(out["fim"].cpu().numpy(), g["fim"])
        d = {k: np.abs(out[k].cpu().numpy() - g[k]).max() for k in ("wim", "cond", "T")}
        d["tsf_inputs"] = np.abs(out["tsf_inputs"].cpu().numpy() - g["tsf_inputs_ac%d" % int(ac)]).max()
        print("align_corners=%d vs reference utils/nmr.py: %s" % (ac, d))
        assert max(d.values()) < TOL
        assert np.array_equal(out["f2verts"][:, ::7].cpu().numpy(), g["f2verts_sub"])
    # default convention = torch 1.2 (align_corners=True)
    dflt = K.correspond(cam[1:].contiguous(), verts[1:].contiguous(), r.faces, 64, r.map_fn, p2v.contiguous(), src_img)
    assert np.abs(dflt["tsf_inputs"].cpu().numpy() - g["tsf_inputs_ac1"]).max() < TOL

    # the standalone mirror methods
    T = r.cal_bc_transform(p2v.expand(2, -1, -1, -1), out["fim"], out["wim"])
    assert np.abs(T.cpu().numpy() - g["T"]).max() < TOL
    T_vis = r.cal_bc_transform(vis.expand(2, -1, -1, -1), out["fim"], out["wim"])
    assert np.abs(T_vis.cpu().numpy() - g["T_vis"]).max() < TOL
    assert np.array_equal(r.encode_front_fim(out["fim"], front_fn=True).cpu().numpy(), g["front"])
    assert np.array_equal(r.encode_front_fim(out["fim"], front_fn=False).cpu().numpy(), g["back"])
