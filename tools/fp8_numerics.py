"""CPU numerics experiment: can the two low-order products of the fp16 hi/lo split run in fp8?

Emulates  x_hi*w_hi (fp16 operands)  +  qa(x)*qb(w_lo)  +  qc(x_lo)*qd(w)  with fp32 accumulation through the
oracle generator and reports the end-to-end error against plain fp32.  Test infrastructure only.
"""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from impersonator_b200 import synthetic as S
from impersonator_b200.generator import ImpersonatorGenerator
from oracle import generator_ref as G

torch.set_grad_enabled(False)
torch.set_num_threads(32)

E5, E4 = torch.float8_e5m2, torch.float8_e4m3fn


def q(t, dt, scale=1.0):
    lim = 57344.0 if dt is E5 else 448.0
    return (t * scale).clamp(-lim, lim).to(dt).float() / scale


def split(t):
    hi = t.half().float()
    lo = (t - hi).half().float()
    return hi, lo


MODE = {"name": "fp32"}


def conv_emul(fn, x, w, **kw):
    m = MODE["name"]
    if m == "fp32" or w.shape[-1] == 7 and w.shape[0] <= 3:
        return fn(x, w, **kw)
    xh, xl = split(x)
    wh, wl = split(w)
    if m == "fp16":
        return fn(xh, wh, **kw)
    if m == "x3":
        return fn(xh, wh, **kw) + fn(xh, wl, **kw) + fn(xl, wh, **kw)
    if m == "x2w":      # x exact-ish, w rounded
        return fn(xh, wh, **kw) + fn(xl, wh, **kw)
    da, db, dc, dd, s, t = MODE["cfg"]
    if "scales" in MODE:
        sa, sb, sc, sd_ = MODE["scales"]
        t2 = fn(q(xh, da, sa), q(wl, db, sb), **kw)
        t3 = fn(q(xl, dc, sc), q(wh, dd, sd_), **kw)
    else:
        t2 = fn(q(xh, da, 2.0 ** -s), q(wl, db, 2.0 ** s), **kw)
        t3 = fn(q(xl, dc, 2.0 ** t), q(wh, dd, 2.0 ** -t), **kw)
    return fn(xh, wh, **kw) + t2 + t3


_c2, _ct = F.conv2d, F.conv_transpose2d


import types
FP = types.SimpleNamespace(**{k: getattr(F, k) for k in dir(F) if not k.startswith("__")})
FP.conv2d = lambda x, w, **kw: conv_emul(_c2, x, w, **kw)
FP.conv_transpose2d = lambda x, w, **kw: conv_emul(_ct, x, w, **kw)
G.F = FP


n = ImpersonatorGenerator(bg_dim=4, src_dim=6, tsf_dim=6, repeat_num=6)
sd = S.fill_state_dict(n.state_dict(), seed=0)
inp = S.synthetic_generator_inputs(1, 256, seed=21)


def run():
    e, r = G.encode_src(inp["src"], sd)
    return G.inference(e, r, inp["tsf"], inp["T"], sd)


img0, mask0 = run()
cfgs = [("x3", None),
        ("e4m3: x*1, wlo*2^15 | xlo*2^12, w*2^3", "e4s"),
        ("e5m2 same scales", "e5s"),
        ]
import sys
for seed in (21, 33):
    inp = S.synthetic_generator_inputs(1, 256, seed=seed)
    MODE["name"] = "fp32"
    img0, mask0 = run()
    for name, cfg in cfgs:
        if cfg is None:
            MODE["name"] = name
        else:
            MODE["name"] = "fp8"
            dt = E4 if cfg == "e4s" else E5
            # conv_emul reads (da, db, dc, dd, s, t): term2 = q(xh, da, 2^-s) q(wl, db, 2^s); term3 = q(xl, dc, 2^t) q(wh, dd, 2^-t)
            MODE["cfg"] = (dt, dt, dt, dt, 0, 0)
            MODE["scales"] = (1.0, 2.0 ** 15, 2.0 ** 12, 2.0 ** 3)
        img, mask = run()
        print("seed %d %-42s img %.3e mask %.3e" % (seed, name, (img - img0).abs().max().item(), (mask - mask0).abs().max().item()), flush=True)
