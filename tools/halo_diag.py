"""Diagnostic: which descriptor encoding makes kx-shifted (non-1024B-aligned) UMMA windows work."""
import os
import subprocess
import sys

CASE = r'''
import sys, torch, torch.nn.functional as F
sys.path.insert(0, ".")
from impersonator_b200 import kernels as K
torch.manual_seed(0)
dev = torch.device("cuda")
def run(kh, kw, cin=64, cout=64, h=32, w=16):
    x = torch.randn(1, cin, h, w); wt = torch.randn(cout, cin, kh, kw) * 0.05
    ref = F.conv2d(x, wt, padding=(kh // 2, kw // 2))
    xs = K.nchw_to_nhwc_split(x.to(dev), split=False)
    ws = K.pack_conv_weight(wt.to(dev), split=False)
    d = K.make_conv_desc(1, h, w, cin, cout, kh, kw, stride=1, pad=kh // 2, split=False, halo=True)
    d.h_out, d.w_out = h, w
    out = torch.full((1, h, w, cout), float("nan"), device=dev)
    K.ConvPlan(d, xs, None, ws, out, None).run()
    torch.cuda.synchronize()
    got = out.permute(0, 3, 1, 2).cpu()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    # per-tap probe: which single taps are right?  (weights of one tap only)
    return err
for (kh, kw) in ((1, 1), (3, 1), (1, 3), (3, 3), (1, 7)):
    print("bo=%s  kernel %dx%d  rel err %.3e" % (__import__("os").environ.get("LWB_HALO_BO"), kh, kw, run(kh, kw)))
# single-tap probes for 1x3: only tap kx active
for kx in range(3):
    x = torch.randn(1, 64, 32, 16); wt = torch.zeros(64, 64, 1, 3); wt[:, :, 0, kx] = torch.randn(64, 64) * 0.05
    ref = F.conv2d(x, wt, padding=(0, 1))
    xs = K.nchw_to_nhwc_split(x.to(dev), split=False); ws = K.pack_conv_weight(wt.to(dev), split=False)
    d = K.make_conv_desc(1, 32, 16, 64, 64, 1, 3, stride=1, pad=0, split=False, halo=True); d.h_out, d.w_out = 32, 16
    out = torch.full((1, 32, 16, 64), float("nan"), device=dev)
    K.ConvPlan(d, xs, None, ws, out, None).run(); torch.cuda.synchronize()
    got = out.permute(0, 3, 1, 2).cpu()
    e = (got - ref).abs()
    # does the result match a different shift?  correlate with refs at other shifts
    best = []
    for s in range(-3, 4):
        wt2 = torch.zeros(64, 64, 1, 7); wt2[:, :, 0, 3 + s] = wt[:, :, 0, kx]
        r2 = F.conv2d(x, wt2, padding=(0, 3))
        best.append(((got - r2).abs().max().item() / r2.abs().max().item(), s))
    print("  1x3 tap kx=%d: rel err %.3e ; err vs other shifts (err, shift): %s ; rows of 8 px wrong-pattern: %s"
          % (kx, e.max().item() / ref.abs().max().item(), sorted(best)[:2], (e.amax(dim=(0, 1, 2)) > 1e-2).int().tolist()))
'''
for bo in ("1", "0", "2"):
    env = dict(os.environ, LWB_HALO_BO=bo)
    r = subprocess.run([sys.executable, "-c", CASE], env=env, capture_output=True, text=True, timeout=300)
    print(r.stdout[-3000:], r.stderr[-1500:] if r.returncode else "")
