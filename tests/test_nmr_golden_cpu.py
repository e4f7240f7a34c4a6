"""The geometry-glue oracle (oracle/nmr_ref.py) against tests/golden/nmr.npz -- the outputs of the REFERENCE's own
utils/nmr.py methods (render_fim_wim, encode_fim, encode_front_fim, get_vis_f2pts, cal_bc_transform), produced by
tests/golden/make_nmr_golden.py in the build container.  Exact: same torch ops, same rasterizer oracle."""
import os

import numpy as np
import torch

from impersonator_b200 import synthetic as S
from oracle import nmr_ref

GOLD = os.path.join(os.path.dirname(__file__), "golden", "nmr.npz")


def _inputs():
    g = np.load(GOLD)
    v, f = S.uv_sphere()
    cam, verts = S.synthetic_frames(3, seed=41, base_verts=v)
    assert np.array_equal(cam.numpy(), g["cam"])
    return g, f, cam, verts, S.synthetic_tables(), S.synthetic_source(64)


def test_oracle_equals_reference_nmr_methods():
    g, f, cam, verts, tabs, src_img = _inputs()
    s_f2v, s_fim, s_wim = nmr_ref.render_fim_wim(cam[:1], verts[:1], f, 64)
    assert np.array_equal(s_fim.numpy(), g["src_fim"]) and np.array_equal(s_wim.numpy(), g["src_wim"])
    assert np.array_equal(s_f2v[:, ::7].numpy(), g["src_f2verts_sub"])
    assert np.array_equal(nmr_ref.encode_fim(s_fim, tabs["map_fn"]).numpy(), g["src_cond"])
    p2v = nmr_ref.src_p2verts(s_f2v)
    vis = nmr_ref.get_vis_f2pts(p2v, s_fim)
    assert np.array_equal(vis[:, ::7].numpy(), g["vis_sub"])
    assert np.array_equal((vis[0, :, 0, 0] != -2).nonzero()[:, 0].numpy(), g["vis_ids"])
    for ac in (True, False):
        c = nmr_ref.correspond(cam[1:], verts[1:], f, tabs["map_fn"], p2v, src_img, 64, align_corners=ac)
        assert np.array_equal(c["tsf_inputs"].numpy(), g["tsf_inputs_ac%d" % int(ac)])
    assert np.array_equal(c["fim"].numpy(), g["fim"]) and np.array_equal(c["wim"].numpy(), g["wim"])
    assert np.array_equal(c["cond"].numpy(), g["cond"]) and np.array_equal(c["T"].numpy(), g["T"])
    assert np.array_equal(c["f2verts"][:, ::7].numpy(), g["f2verts_sub"])
    assert np.array_equal(nmr_ref.cal_bc_transform(vis.expand(2, -1, -1, -1), c["fim"], c["wim"], 64).numpy(), g["T_vis"])
    assert np.array_equal(nmr_ref.encode_fim(c["fim"], tabs["front_map_fn"]).numpy(), g["front"])
    assert np.array_equal(nmr_ref.encode_fim(c["fim"], tabs["back_map_fn"]).numpy(), g["back"])
    # sanity of the golden itself: the silhouettes are non-trivial and T is -2 exactly off the body
    assert 0.05 < (g["fim"] >= 0).mean() < 0.9
    assert np.all(g["T"][g["fim"] < 0] == -2.0)
