"""CPU model behind the head-range flag (k_heads, bit 4): how the end-to-end error of the fp16f8 operand split grows with
the scale of the output heads' pre-activations.

Emulates the engine's arithmetic through the oracle generator -- per conv:  fp16(x) * fp16(w)  (fp32 accumulate)
+ e4m3(x/16) * e4m3(w_lo * 2^(E+4)) + e4m3(x_lo * 2^10) * e4m3(w * 2^(E-10))  with the per-layer power-of-two weight
exponent E of kernels.weight_exponent, the heads included (they run on the same engine) -- on the benchmark's synthetic
weights with the two head filters scaled by s, and prints max |pre-activation| next to the max-abs error of the colour
and mask outputs against plain fp32.  Test infrastructure only (imports oracle/).

    python tools/head_range_model.py            # ~2 min on 8 threads
"""
import math
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from impersonator_b200 import synthetic as S
from impersonator_b200.generator import ImpersonatorGenerator
from oracle import generator_ref as G

torch.set_grad_enabled(False)
E4 = torch.float8_e4m3fn
MODE = {"name": "fp32"}


def q8(t, scale):
    return (t * scale).clamp(-448.0, 448.0).to(E4).float() / scale


def split(t):
    hi = t.half().float()
    return hi, (t - hi).half().float()


def conv_emul(fn, x, w, **kw):
    if MODE["name"] == "fp32":
        return fn(x, w, **kw)
    xh, xl = split(x)
    wh, wl = split(w)
    if MODE["name"] == "fp16x3":
        return fn(xh, wh, **kw) + fn(xh, wl, **kw) + fn(xl, wh, **kw)
    if MODE["name"] == "fp16" or tuple(w.shape) in MODE.get("plain", ()):
        return fn(xh, wh, **kw)
    E = 14 - math.floor(math.log2(float(w.abs().max())))          # max|w| * 2^E in [2^14, 2^15)
    out = fn(xh, wh, **kw)
    if MODE["name"] != "fp16f8-no-wlo":
        out = out + fn(q8(x, 2.0 ** -4), q8(wl, 2.0 ** (E + 4)), **kw)
    if MODE["name"] != "fp16f8-no-xlo":
        out = out + fn(q8(xl, 2.0 ** 10), q8(w, 2.0 ** (E - 10)), **kw)
    return out


_c2, _ct = F.conv2d, F.conv_transpose2d
FP = types.SimpleNamespace(**{k: getattr(F, k) for k in dir(F) if not k.startswith("__")})
FP.conv2d = lambda x, w, **kw: conv_emul(_c2, x, w, **kw)
FP.conv_transpose2d = lambda x, w, **kw: conv_emul(_ct, x, w, **kw)
G.F = FP


def preacts(x, sd, p='tsf_model'):
    return (FP.conv2d(x, sd[p + '.img_reg.0.weight'], padding=3), FP.conv2d(x, sd[p + '.attetion_reg.0.weight'], padding=3))


def run(sd, inp):
    e, r = G.encode_src(inp["src"], sd)
    tsf_x = G.unet_encoder(inp["tsf"], sd, 'tsf_model', 0)
    outs = [tsf_x]
    for i in range(1, 4):
        tsf_x = G.unet_encoder(tsf_x, sd, 'tsf_model', i) + G.transform(e[i], inp["T"])
        outs.append(tsf_x)
    Ts = G.resize_trans(e[3], inp["T"])
    for i in range(6):
        tsf_x = G.residual_block(tsf_x, sd, 'tsf_model.resnets.%d' % i) + G.stn(r[i], Ts)
    a_img, a_mask = preacts(G.unet_decode(tsf_x, outs, sd, 'tsf_model'), sd)
    return a_img, a_mask


def main():
    n = ImpersonatorGenerator(bg_dim=4, src_dim=6, tsf_dim=6, repeat_num=6)
    base = S.fill_state_dict(n.state_dict(), seed=0)
    size = int(os.environ.get("SIZE", "256"))
    print("%-6s %-12s %-12s | %-22s | %-22s" % ("scale", "max|a_img|", "max|a_mask|", "fp16f8  img / mask", "fp16x3  img / mask"))
    for s in (1, 2, 4, 8, 16):
        sd = dict(base)
        for k in ('tsf_model.img_reg.0.weight', 'tsf_model.attetion_reg.0.weight'):
            sd[k] = base[k] * s
        worst = {}
        amax = [0.0, 0.0]
        for seed in (21, 33):
            inp = S.synthetic_generator_inputs(1, size, seed=seed)
            MODE["name"] = "fp32"
            a0, m0 = run(sd, inp)
            amax = [max(amax[0], a0.abs().max().item()), max(amax[1], m0.abs().max().item())]
            for mode in ("fp16f8", "fp16x3"):
                MODE["name"] = mode
                a, m = run(sd, inp)
                e = ((torch.tanh(a) - torch.tanh(a0)).abs().max().item(), (torch.sigmoid(m) - torch.sigmoid(m0)).abs().max().item())
                worst[mode] = tuple(max(x, y) for x, y in zip(worst.get(mode, (0, 0)), e))
        print("%-6d %-12.2f %-12.2f | %.2e / %.2e    | %.2e / %.2e" % (s, amax[0], amax[1], worst["fp16f8"][0], worst["fp16f8"][1],
                                                                     worst["fp16x3"][0], worst["fp16x3"][1]), flush=True)


def variants():
    """Why both correction products are needed: the benchmark weights with one of them dropped, and plain fp16."""
    n = ImpersonatorGenerator(bg_dim=4, src_dim=6, tsf_dim=6, repeat_num=6)
    sd = S.fill_state_dict(n.state_dict(), seed=0)
    size = int(os.environ.get("SIZE", "256"))
    inp = S.synthetic_generator_inputs(1, size, seed=21)
    MODE["name"] = "fp32"
    a0, m0 = run(sd, inp)
    for mode in ("fp16x3", "fp16f8", "fp16f8-no-xlo", "fp16f8-no-wlo", "fp16"):
        MODE["name"] = mode
        a, m = run(sd, inp)
        print("%-14s colour %.2e  mask %.2e" % (mode, (torch.tanh(a) - torch.tanh(a0)).abs().max().item(),
                                                (torch.sigmoid(m) - torch.sigmoid(m0)).abs().max().item()), flush=True)


def per_layer():
    """Which layers could run the single fp16 product (no correction MMAs) inside an otherwise fp16f8 network: the
    layers are selected by weight shape (Conv2d [cout,cin,k,k]; ConvTranspose2d [cin,cout,k,k])."""
    n = ImpersonatorGenerator(bg_dim=4, src_dim=6, tsf_dim=6, repeat_num=6)
    sd = S.fill_state_dict(n.state_dict(), seed=0)
    size = int(os.environ.get("SIZE", "256"))
    inp = S.synthetic_generator_inputs(1, size, seed=21)
    MODE["name"] = "fp32"
    a0, m0 = run(sd, inp)
    MODE["name"] = "fp16f8"
    sets = [("none (fp16f8 everywhere)", ()),
            ("heads 64->3/1 7x7", ((3, 64, 7, 7), (1, 64, 7, 7))),
            ("stem 6->64 7x7", ((64, 6, 7, 7),)),
            ("skipper 128->64 @256", ((64, 128, 3, 3),)),
            ("convT 128->64", ((128, 64, 3, 3),)),
            ("enc 64->128 s2", ((128, 64, 3, 3),)),
            ("all of the above", ((3, 64, 7, 7), (1, 64, 7, 7), (64, 6, 7, 7), (64, 128, 3, 3), (128, 64, 3, 3))),
            ("twelve 512->512", ((512, 512, 3, 3),))]
    for name, shapes in sets:
        MODE["plain"] = shapes
        a, m = run(sd, inp)
        print("%-28s colour %.2e  mask %.2e" % (name, (torch.tanh(a) - torch.tanh(a0)).abs().max().item(),
                                                (torch.sigmoid(m) - torch.sigmoid(m0)).abs().max().item()), flush=True)
    MODE.pop("plain", None)


if __name__ == "__main__":
    per_layer() if "--per-layer" in sys.argv else variants() if "--variants" in sys.argv else main()
