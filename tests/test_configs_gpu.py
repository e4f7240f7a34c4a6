"""BASELINE.json parity cases that are not the bench workload: configs[1] (batch 8 raster + LWB +
generator forward) and configs[4] (512x512 high-res variant), checked against the oracle."""
import numpy as np
import pytest
import torch

from impersonator_b200 import kernels as K
from impersonator_b200 import synthetic as S
from impersonator_b200.generator import ImpersonatorGenerator
from oracle import generator_ref as G, nmr_ref

pytestmark = pytest.mark.gpu


def _run(cuda, batch, size, oracle_frames):
    torch.set_grad_enabled(False)
    v, f = S.uv_sphere()
    cam, verts = S.synthetic_frames(batch + 1, seed=77, base_verts=v)
    tabs = S.synthetic_tables()
    src_img = S.synthetic_source(size)
    net = ImpersonatorGenerator(bg_dim=4, src_dim=6, tsf_dim=6, repeat_num=6)
    sd = S.fill_state_dict(net.state_dict(), seed=0)
    net.load_state_dict(sd)
    net = net.to(cuda).eval()
    f2v, sfim, _ = nmr_ref.render_fim_wim(cam[:1], verts[:1], f, size)
    p2v = nmr_ref.src_p2verts(f2v)
    src_inputs = torch.cat([src_img, nmr_ref.encode_fim(sfim, tabs["map_fn"])], dim=1)
    bg = torch.zeros(1, 3, size, size)
    out = K.correspond(cam[1:].to(cuda).contiguous(), verts[1:].to(cuda).contiguous(), f.to(cuda), size,
                       tabs["map_fn"].to(cuda), p2v.to(cuda).contiguous(), src_img.to(cuda))
    enc, res = net.encode_src(src_inputs.to(cuda))
    _, _, pred = net.inference(enc, res, out["tsf_inputs"], out["T"], bg=bg.to(cuda))
    torch.cuda.synchronize()
    sel = list(range(1, 1 + oracle_frames))
    ref = nmr_ref.correspond(cam[sel], verts[sel], f, tabs["map_fn"], p2v, src_img, size)
    assert int((out["fim"][:oracle_frames].cpu() != ref["fim"]).sum()) == 0
    assert (out["T"][:oracle_frames].cpu() - ref["T"]).abs().max().item() < 1e-5
    feats = G.encode_src(src_inputs, sd)
    ref_pred, _, _ = G.imitator_forward(bg, feats, ref["tsf_inputs"], ref["T"], sd)
    d = (pred[:oracle_frames].cpu() - ref_pred).abs().max().item()
    print("batch %d @%d: pred max-abs vs oracle %.3e (first %d frames)" % (batch, size, d, oracle_frames))
    assert d < 1e-3
    # the frames the oracle did not evaluate must at least be distinct, finite results
    assert torch.isfinite(pred).all()


def test_config1_batch8_256(cuda):
    _run(cuda, 8, 256, 2)


def test_config4_512_highres(cuda):
    _run(cuda, 2, 512, 1)
