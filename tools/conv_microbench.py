"""Micro-benchmark of single conv layers (CUDA events): stats on/off, halo on/off, N tile."""
import sys
import torch
sys.path.insert(0, ".")
from impersonator_b200 import kernels as K

dev = torch.device("cuda")
torch.manual_seed(0)


def bench(name, n, cin, cout, h, k, stats=True, halo=False, n_tile=0, split=True, rowk=False, cin1=0, reps=20):
    if rowk:
        x = (torch.randn(n, h + 6, h + 8, 8, device=dev).half(), torch.randn(n, h + 6, h + 8, 8, device=dev).half())
        w = (torch.randn(7, cout, 64, device=dev).half(), torch.randn(7, cout, 64, device=dev).half())
        d = K.make_conv_desc(n, h, h, 8, cout, 7, 7, stride=1, pad=3, split=split, rowk=True, row_pitch=h + 8, halo=halo, n_tile=n_tile)
        x1 = None
    else:
        x = (torch.randn(n, h, h, cin, device=dev).half(), torch.randn(n, h, h, cin, device=dev).half())
        x1 = (torch.randn(n, h, h, cin1, device=dev).half(), torch.randn(n, h, h, cin1, device=dev).half()) if cin1 else None
        w = (torch.randn(k * k, cout, cin + cin1, device=dev).half(), torch.randn(k * k, cout, cin + cin1, device=dev).half())
        d = K.make_conv_desc(n, h, h, cin, cout, k, k, stride=1, pad=k // 2, cin1=cin1, split=split, halo=halo, n_tile=n_tile)
    out = torch.empty((n, h, h, cout), device=dev)
    st = torch.zeros((n, cout, 2), dtype=torch.float64, device=dev) if stats else None
    plan = K.ConvPlan(d, x, x1, w, out, st)
    for _ in range(3):
        plan.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        plan.run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-44s stats=%d halo=%d ntile=%3d : %7.3f ms  %7.1f TF/s algorithmic" % (name, stats, halo, n_tile or -1, ms, plan.flops / ms / 1e9))


for stats in (True,):
    for halo in (False, True):
        bench("stem rowk 8->64 @256 B16", 16, 8, 64, 256, 7, stats=stats, halo=halo, rowk=True)
        bench("skipper 64+64->64 @256 B16", 16, 64, 64, 256, 3, stats=stats, halo=halo, cin1=64)
        bench("skipper 128+128->128 @128 B16", 16, 128, 128, 128, 3, stats=stats, halo=halo, cin1=128)
        bench("res 512->512 @32 B16", 16, 512, 512, 32, 3, stats=stats, halo=halo)
bench("res 512->512 @32 B16 fast", 16, 512, 512, 32, 3, stats=True, split=False)
bench("res 512->512 @32 B16 fast n128", 16, 512, 512, 32, 3, stats=True, split=False, n_tile=128)
bench("res 512->512 @32 B16 split n256", 16, 512, 512, 32, 3, stats=True, n_tile=256)
bench("skipper 64+64->64 @256 B16 fast", 16, 64, 64, 256, 3, stats=True, cin1=64, split=False)
bench("stem rowk 8->64 @256 B16 halo(resident B)", 16, 8, 64, 256, 7, stats=True, halo=True, rowk=True)
bench("stem rowk 8->64 @256 B16 halo fast", 16, 8, 64, 256, 7, stats=True, halo=True, rowk=True, split=False)
bench("stem rowk 8->64 @256 B16 fast", 16, 8, 64, 256, 7, stats=True, halo=False, rowk=True, split=False)
