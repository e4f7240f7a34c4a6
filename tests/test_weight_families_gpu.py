"""Parity under weight families other than N(0, 0.02) (SURVEY.md 8c / VERDICT r1): PyTorch's default init, a
heavy-tailed "trained-like" set (per-layer conv std up to 0.5, outliers |w| > 2, InstanceNorm gains up to 4), and a
range-stress set whose residual stream exceeds the e4m3 correction range.  The default fp16f8 mode packs every layer
with its own power-of-two weight scale (no |w| < 2 restriction) and reports out-of-range activations instead of
saturating silently.  Oracle: oracle/generator_ref.py (== the reference modules, tests/golden/make_generator_golden.py)."""
import numpy as np
import pytest
import torch

from impersonator_b200 import kernels as K
from impersonator_b200 import synthetic as S
from impersonator_b200.generator import ImpersonatorGenerator
from oracle import generator_ref as G

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _run(cuda, sd, B=2, seed=21):
    n = ImpersonatorGenerator(bg_dim=4, src_dim=6, tsf_dim=6, repeat_num=6)
    n.load_state_dict(sd)
    n = n.to(cuda).eval()
    inp = S.synthetic_generator_inputs(B, 256, seed=seed)
    enc, res = n.encode_src(inp["src"].to(cuda))
    img, mask = n.inference(enc, res, inp["tsf"].to(cuda), inp["T"].to(cuda))
    status = n.range_status()
    e_o, r_o = G.encode_src(inp["src"], sd)
    img_o, mask_o = G.inference(e_o, r_o, inp["tsf"], inp["T"], sd)
    d = {"img": (img.cpu() - img_o).abs().max().item(), "mask": (mask.cpu() - mask_o).abs().max().item(),
         "enc3_rel": ((enc[3].cpu() - e_o[3]).abs().max() / e_o[3].abs().max()).item(),
         "res5_rel": ((res[5].cpu() - r_o[5]).abs().max() / r_o[5].abs().max()).item(),
         "res5_absmax": r_o[5].abs().max().item()}
    return n, d, status, (inp, sd, img_o, mask_o)


def test_pytorch_default_init(cuda):
    torch.set_grad_enabled(False)
    torch.manual_seed(0)
    sd = {k: v.clone() for k, v in ImpersonatorGenerator(bg_dim=4, src_dim=6, tsf_dim=6, repeat_num=6).state_dict().items()}
    _, d, status, _ = _run(cuda, sd)
    print("default init:", d, "range", status)
    assert d["img"] < TOL and d["mask"] < TOL and d["enc3_rel"] < TOL and d["res5_rel"] < TOL and not (status & 3)


def trained_like(seed=0, gain=4.0):
    """Student-t conv weights with a per-layer std drawn log-uniformly from [0.02, 0.5] and a few |w| > 2 outliers;
    InstanceNorm gains log-uniform in [1/gain, gain], biases N(0, 0.5).  The two 7x7 heads (no normalisation behind them)
    get std in [0.004, 0.03] so that the pre-tanh / pre-sigmoid values stay O(1..10) as in a trained network -- with std 0.5
    they would be ~50 and a 1e-4 relative error of ANY fp32 implementation already exceeds 1e-3 on the unsaturated pixels."""
    tmpl = ImpersonatorGenerator(bg_dim=4, src_dim=6, tsf_dim=6, repeat_num=6).state_dict()
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, t in tmpl.items():
        if t.dim() == 4:
            std = float(0.02 * (25.0 ** torch.rand(1, generator=g)))
            if "img_reg" in k or "attetion_reg" in k:
                std = float(0.004 * (7.5 ** torch.rand(1, generator=g)))
            w = torch.distributions.StudentT(4.0).sample(t.shape) * std / 1.414
            idx = torch.randint(0, w.numel(), (8,), generator=g)
            w.view(-1)[idx] = torch.sign(w.view(-1)[idx]) * (2.0 + 3.0 * torch.rand(8, generator=g))
            sd[k] = w.float()
        elif k.endswith("weight"):
            sd[k] = (gain ** (2 * torch.rand(t.shape, generator=g) - 1)).float()
        else:
            sd[k] = (0.5 * torch.randn(t.shape, generator=g)).float()
    return sd


def test_trained_like_heavy_tailed_weights(cuda):
    torch.set_grad_enabled(False)
    torch.manual_seed(1)
    sd = trained_like(seed=0)
    assert max(v.abs().max().item() for v in sd.values() if v.dim() == 4) > 2.0        # the old packing would have raised
    n, d, status, (inp, sd, img_o, mask_o) = _run(cuda, sd)
    print("trained-like (fp16f8):", d, "range", status)
    n.set_precision("fp16x3")
    enc, res = n.encode_src(inp["src"].to(cuda))
    img, mask = n.inference(enc, res, inp["tsf"].to(cuda), inp["T"].to(cuda))
    d3 = {"img": (img.cpu() - img_o).abs().max().item(), "mask": (mask.cpu() - mask_o).abs().max().item()}
    print("trained-like (fp16x3):", d3)
    # Features are always precise; on the PIXELS fp16f8's ~1e-4 relative end-to-end precision is multiplied by the scale of
    # the head pre-activations: either the 1e-3 bar holds, or the run is flagged (bit 2: |pre-activation| >= 8) so that the
    # caller (Imitator.inference does it automatically) switches to fp16x3 -- which must then meet the bar.
    assert d["enc3_rel"] < TOL and d["res5_rel"] < TOL and not (status & 3)
    assert (d["img"] < TOL and d["mask"] < TOL) or (status & 4), "error above the bar and not reported"
    assert d3["img"] < TOL and d3["mask"] < TOL


def test_range_stress_is_reported_and_fp16x3_recovers(cuda):
    """InstanceNorm gains of ~300 push the residual stream past 1024: fp16f8 keeps running (graceful degradation of the
    e4m3 correction terms) but raises bit 0 of the range flag; pinning the network to fp16x3 restores full precision."""
    torch.set_grad_enabled(False)
    torch.manual_seed(2)
    sd = S.fill_state_dict(ImpersonatorGenerator(bg_dim=4, src_dim=6, tsf_dim=6, repeat_num=6).state_dict(), seed=0)
    for k in list(sd):
        if ".resnets." in k and k.endswith("main.4.weight"):
            sd[k] = sd[k] * 300.0
    n, d, status, (inp, sd, img_o, mask_o) = _run(cuda, sd)
    print("range stress (fp16f8):", d, "range", status)
    assert d["res5_absmax"] > 1024.0
    assert status & 1, "activations beyond the e4m3 correction range must be reported"
    assert not (status & 2)
    n.set_precision("fp16x3")
    enc, res = n.encode_src(inp["src"].to(cuda))
    img, mask = n.inference(enc, res, inp["tsf"].to(cuda), inp["T"].to(cuda))
    d3 = {"img": (img.cpu() - img_o).abs().max().item(), "mask": (mask.cpu() - mask_o).abs().max().item()}
    print("range stress (fp16x3):", d3)
    assert d3["img"] < TOL and d3["mask"] < TOL
