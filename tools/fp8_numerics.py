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
cfgs = [("fp16", None), ("x3", None), ("x2w", None),
        ("e5 all s8", (E5, E5, E5, E5, 8, 0)),
        ("e4x e5wl s8 | e5xl e4w", (E4, E5, E5, E4, 8, 0)),
        ("e4x(s4) e4wl(s12)| e5xl e4w(t-4)", (E4, E4, E5, E4, 0, 0)),
        ("e5 all s0", (E5, E5, E5, E5, 0, 0)),
        ]
for name, cfg in cfgs:
    MODE["name"] = name if cfg is None else "fp8"
    MODE["cfg"] = cfg
    img, mask = run()
    print("%-40s img %.3e mask %.3e" % (name, (img - img0).abs().max().item(), (mask - mask0).abs().max().item()), flush=True)
