"""Generates tests/golden/generator.npz -- run ONLY in the build container (needs /root/reference).

Imports the reference's own networks/generator.py (pure torch.nn; ipdb/h5py stubbed), loads the
deterministic weights of impersonator_b200.synthetic.fill_state_dict(seed=0) into it, runs
  ImpersonatorGenerator.forward      (networks/generator.py:204-211)   B=1   (BASELINE config 1)
  encode_src + inference             (:213-214, :277-301)              B=2
  swap                               (:245-275)                        B=1, two sources
on impersonator_b200.synthetic.synthetic_generator_inputs and stores strided slices of every output;
plus the two BASELINE sizes no other golden covers -> generator_big.npz:
  encode_src + inference             B=16 @256x256 (configs[2]) and B=8 @512x512 (configs[4]).

grid_sample convention: the reference calls F.grid_sample without ``align_corners`` (networks/generator.py:313)
under its pinned torch==1.2.0, where that means True.  The installed torch 2.11 would silently use False
for the same call, so the reference modules are run under ``torch12_grid_sample()`` (the flag-less call ->
align_corners=True; nothing else is touched) for the default keys, and once more unpatched for the
``ac0_*`` keys that pin the opt-in LWB_ALIGN_CORNERS=0 mode.
It also checks oracle/generator_ref.py (the functional restatement) against the reference modules
on the full tensors, so the restatement is pinned to the reference here, and the slices pin both
on the GPU box (where /root/reference does not exist).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
for m in ("ipdb", "h5py"):
    sys.modules.setdefault(m, types.ModuleType(m))
sys.path.insert(0, "/root/reference")

from networks.generator import ImpersonatorGenerator          # noqa: E402  (the reference)
from impersonator_b200 import synthetic                       # noqa: E402
from oracle import generator_ref as G                         # noqa: E402


def sl(t):
    return t[:, :, 3::8, 5::8].contiguous().numpy()


class torch12_grid_sample(object):
    """F.grid_sample without the flag behaves as in torch 1.2 (align_corners=True) inside the block."""

    def __enter__(self):
        import torch.nn.functional as F
        self.F, self.orig = F, F.grid_sample

        def gs(input, grid, mode='bilinear', padding_mode='zeros', align_corners=None):
            return self.orig(input, grid, mode=mode, padding_mode=padding_mode,
                             align_corners=True if align_corners is None else align_corners)
        F.grid_sample = gs
        return self

    def __exit__(self, *a):
        self.F.grid_sample = self.orig
        return False


def main():
    torch.set_grad_enabled(False)
    net = ImpersonatorGenerator(bg_dim=4, src_dim=6, tsf_dim=6, repeat_num=6).eval()
    sd = synthetic.fill_state_dict(net.state_dict(), seed=0)
    net.load_state_dict(sd, strict=True)
    keys = sorted(sd.keys())
    out = {"keys": np.array(keys), "shapes": np.array([str(tuple(sd[k].shape)) for k in keys])}

    # ---- opt-in convention (installed torch, align_corners=False): a few keys only
    inp = synthetic.synthetic_generator_inputs(1, 256, seed=11)
    ref0 = net(inp["bg"], inp["src"], inp["tsf"], inp["T"])
    mine0 = G.forward(inp["bg"], inp["src"], inp["tsf"], inp["T"], sd, align_corners=False)
    for name, a, b in zip(("img_bg", "src_img", "src_mask", "tsf_img", "tsf_mask"), ref0, mine0):
        assert (a - b).abs().max().item() < 1e-5
        out["ac0_fwd_" + name] = sl(a)
    with torch12_grid_sample():
        fill(net, sd, out)
    np.savez_compressed(os.path.join(HERE, "generator.npz"), **out)
    print("wrote generator.npz", {k: v.shape for k, v in out.items() if k not in ("keys", "shapes")})
    with torch12_grid_sample():
        big(net, sd)


def big(net, sd):
    """configs[2] / configs[4] sizes through the reference modules (B=16 @256, B=8 @512)."""
    out = {}
    for tag, B, size, seed, step in (("b16_256", 16, 256, 61, 16), ("b8_512", 8, 512, 71, 32)):
        inp = synthetic.synthetic_generator_inputs(B, size, seed=seed)
        enc, res = net.encode_src(inp["src"])
        img, mask = net.inference([e.expand(B, -1, -1, -1) for e in enc], [e.expand(B, -1, -1, -1) for e in res],
                                  inp["tsf"], inp["T"])
        out[tag + "_img"] = img[:, :, 3::step, 5::step].contiguous().numpy()
        out[tag + "_mask"] = mask[:, :, 3::step, 5::step].contiguous().numpy()
        out[tag + "_img_mean"] = img.mean(dim=(1, 2, 3)).numpy()          # every pixel of every frame contributes
        out[tag + "_img_absmean"] = img.abs().mean(dim=(1, 2, 3)).numpy()
        print(tag, "reference modules done", out[tag + "_img"].shape)
    np.savez_compressed(os.path.join(HERE, "generator_big.npz"), **out)
    print("wrote generator_big.npz")


def fill(net, sd, out):
    inp = synthetic.synthetic_generator_inputs(1, 256, seed=11)
    ref = net(inp["bg"], inp["src"], inp["tsf"], inp["T"])
    mine = G.forward(inp["bg"], inp["src"], inp["tsf"], inp["T"], sd)
    for name, a, b in zip(("img_bg", "src_img", "src_mask", "tsf_img", "tsf_mask"), ref, mine):
        d = (a - b).abs().max().item()
        print("forward  %-9s restatement-vs-reference max-abs %.3g" % (name, d))
        assert d < 1e-5
        out["fwd_" + name] = sl(a)

    inp2 = synthetic.synthetic_generator_inputs(2, 256, seed=21)
    enc, res = net.encode_src(inp2["src"])
    enc2 = [e.expand(2, -1, -1, -1) for e in enc]
    res2 = [e.expand(2, -1, -1, -1) for e in res]
    img, mask = net.inference(enc2, res2, inp2["tsf"], inp2["T"])
    e_m, r_m = G.encode_src(inp2["src"], sd)
    img_m, mask_m = G.inference(e_m, r_m, inp2["tsf"], inp2["T"], sd)
    for name, a, b in (("tsf_img", img, img_m), ("tsf_mask", mask, mask_m), ("enc3", enc[3], e_m[3]),
                       ("res5", res[5], r_m[5])):
        d = (a - b).abs().max().item()
        print("inference %-9s restatement-vs-reference max-abs %.3g" % (name, d))
        assert d < 1e-5
    out["inf_tsf_img"] = sl(img)
    out["inf_tsf_mask"] = sl(mask)
    out["inf_enc3"] = enc[3][:, ::16, ::4, ::4].contiguous().numpy()
    out["inf_res5"] = res[5][:, ::16, ::4, ::4].contiguous().numpy()
    # swap (appearance transfer, networks/generator.py:245-275): two sources, two flows
    a = synthetic.synthetic_generator_inputs(1, 256, seed=31)
    b = synthetic.synthetic_generator_inputs(1, 256, seed=41)
    e12, r12 = net.encode_src(a["src"])
    e21, r21 = net.encode_src(b["src"])
    s_img, s_mask = net.swap(a["tsf"], e12, e21, r12, r21, a["T"], b["T"])
    o12, q12 = G.encode_src(a["src"], sd)
    o21, q21 = G.encode_src(b["src"], sd)
    m_img, m_mask = G.swap(a["tsf"], o12, o21, q12, q21, a["T"], b["T"], sd)
    for name, x, y in (("swap_img", s_img, m_img), ("swap_mask", s_mask, m_mask)):
        d = (x - y).abs().max().item()
        print("swap %-9s restatement-vs-reference max-abs %.3g" % (name, d))
        assert d < 1e-5
        out[name] = sl(x)


if __name__ == "__main__":
    main()
