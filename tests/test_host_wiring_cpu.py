"""Host logic of the network mirrors on CPU: the streams (buffer binding, BatchNorm folding, gated-filter stacking, 7x1 head
folding, LWB plumbing, shortcut / post-affine wiring) run with torch stand-ins for the kernel front-ends
(tests/kernel_emulator.py, same contracts as include/lwb_b200.h) and are compared with the oracles.  The kernels themselves
are covered by the -m gpu tests; this file makes sure the orchestration around them is right without a GPU."""
import numpy as np
import torch

import kernel_emulator
from impersonator_b200 import synthetic as S

CPU = torch.device("cpu")


def test_generator_streams_match_oracle(monkeypatch):
    from impersonator_b200.generator import ImpersonatorGenerator
    from oracle import generator_ref as G
    kernel_emulator.install(monkeypatch)
    torch.set_grad_enabled(False)
    n = ImpersonatorGenerator(bg_dim=4, src_dim=6, tsf_dim=6, repeat_num=6).eval()
    sd = S.fill_state_dict(n.state_dict(), seed=0)
    n.load_state_dict(sd)
    inp = S.synthetic_generator_inputs(2, 64, seed=21)
    for tc_heads in ("1", "0"):
        monkeypatch.setenv("LWB_TC_HEADS", tc_heads)
        n._lwb_invalidate()
        enc, res = n.encode_src(inp["src"])
        bg = torch.rand(1, 3, 64, 64) * 2 - 1
        img, mask, pred = n.inference(enc, res, inp["tsf"], inp["T"], bg=bg)
        e_o, r_o = G.encode_src(inp["src"], sd)
        img_o, mask_o = G.inference(e_o, r_o, inp["tsf"], inp["T"], sd)
        d = max((img - img_o).abs().max().item(), (mask - mask_o).abs().max().item(),
                (pred - (mask_o * bg + (1 - mask_o) * img_o)).abs().max().item(), (enc[3] - e_o[3]).abs().max().item())
        assert d < 2e-4, (tc_heads, d)
    outs = n(inp["bg"], inp["src"], inp["tsf"][:1], inp["T"][:1])
    refs = G.forward(inp["bg"], inp["src"], inp["tsf"][:1], inp["T"][:1], sd)
    assert max((a - b).abs().max().item() for a, b in zip(outs, refs)) < 2e-4
    assert n.range_status() == 0
    # range bits are reported for the streams of the most recent pass only: a stale bit of another shape does not leak
    n.inference(enc, res, inp["tsf"], inp["T"])
    stale = next(st for k, st in n.tsf_model._lwb_streams.items() if k[1] == 2)
    n.inference(enc, res, inp["tsf"][:1], inp["T"][:1])
    stale.range_flag.fill_(4)
    assert len(n.tsf_model.range_flags()) == 1 and n.tsf_model.range_status() == 0
    n.inference(enc, res, inp["tsf"], inp["T"])
    assert n.tsf_model.range_status() == 0                         # the pass zeroes its own flag first


def test_hmr_stream_matches_oracle(monkeypatch):
    from impersonator_b200.hmr import HumanModelRecovery, _HmrStream
    from oracle import hmr_ref
    kernel_emulator.install(monkeypatch)
    torch.set_grad_enabled(False)
    net = HumanModelRecovery(smpl_model=S.synthetic_smpl_model(seed=3)).eval()
    sd = S.synthetic_hmr_state(net.state_dict())
    full = dict(net.state_dict())
    full.update(sd)
    net.load_state_dict(full, strict=True)
    x = S.synthetic_hmr_inputs(2)
    theta = _HmrStream(net, 2, CPU, 1).run(x)
    ref = hmr_ref.forward(x, sd)
    d = (theta - ref).abs().max().item()
    assert d < 1e-4, d


def test_inpaintor_stream_matches_oracle(monkeypatch):
    from impersonator_b200.inpaintor import InpaintSANet, _InpaintStream
    from oracle import inpaintor_ref as R
    kernel_emulator.install(monkeypatch)
    torch.set_grad_enabled(False)
    net = InpaintSANet(c_dim=4).eval()
    sd = S.fill_state_dict(net.state_dict(), seed=3, conv_std=0.05)
    net.load_state_dict(sd)
    img = S.synthetic_source(64, seed=5)
    ys, xs = torch.meshgrid(torch.linspace(-1, 1, 64), torch.linspace(-1, 1, 64), indexing="ij")
    mask = (((xs / 0.4) ** 2 + (ys / 0.8) ** 2) < 1).float()[None, None]
    st = _InpaintStream(net, 1, 64, 64, CPU, 1)
    masked = img * (1 - mask) + mask
    coarse = st.run_coarse(torch.cat([masked, mask], dim=1))
    x = st.run_refine(torch.cat([img * (1 - mask) + coarse * mask, mask], dim=1))
    c_o, x_o, _ = R.forward(img, mask, sd)
    d = max((coarse - c_o).abs().max().item(), (x - x_o).abs().max().item())
    assert d < 2e-4, d


def test_captured_step_pins_evicted_streams_and_epoch_bumps():
    """A CUDA graph replays into the buffers of the per-shape streams it was captured with: the LRU cache may evict them,
    the capturing step must keep them alive (graph.pin), and a parameter reload must change the graph key."""
    from impersonator_b200 import generator as G, graph

    class Holder(object):
        pass

    class FakeStream(object):
        def __init__(self, mod, tag):
            self.tag = tag

    mod = Holder()
    keep = []
    graph._OPEN.append(keep)
    try:
        first = G._stream_for(mod, FakeStream, ('a',), 'a')
        assert G._stream_for(mod, FakeStream, ('a',), 'a') is first
    finally:
        graph._OPEN.pop()
    assert keep and all(o is first for o in keep)
    for i in range(12):                                           # push 'a' out of the cache
        G._stream_for(mod, FakeStream, ('k', i), i)
    assert ('a',) not in mod._lwb_streams and len(mod._lwb_streams) <= 8
    assert keep[0] is first and first.tag == 'a'                  # still referenced by the (fake) captured step
    assert G._stream_for(mod, FakeStream, ('a',), 'a') is not first

    e0 = G.weights_epoch()
    net = G.ResNetGenerator(conv_dim=8, c_dim=4, repeat_num=1, k_size=3, n_down=1)
    net.load_state_dict(net.state_dict())
    assert G.weights_epoch() > e0
