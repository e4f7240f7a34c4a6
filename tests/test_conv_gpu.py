"""GPU parity of the conv engine (tcgen05 implicit GEMM) and its glue kernels against plain
torch fp32 ops on CPU (the same ATen ops the reference's nn.Conv2d / ConvTranspose2d /
InstanceNorm2d / grid_sample calls resolve to).  Tolerances: split (parity) mode 2e-4 of the
output scale; single-pass fp16 ("fast") mode 2e-2."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from impersonator_b200 import kernels as K

pytestmark = pytest.mark.gpu


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def report(name, got, ref):
    d = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-12
    idx = np.unravel_index(int(d.argmax()), d.shape)
    print("%s: max-abs %.3e (ref scale %.3e, rel %.3e) at %s; mean-abs %.3e; nonfinite %d"
          % (name, d.max().item(), scale, d.max().item() / scale, idx, d.mean().item(),
             int((~torch.isfinite(got)).sum())))
    return d.max().item() / scale


def run_conv(cuda, x, w, stride=1, pad=1, dil=1, transposed=False, split=True, x1=None, n_tile=0, stats=True, halo=False,
             pad_w=None):
    """x [n,c,h,w] fp32 CPU, w OIHW (or IOHW when transposed) -> NCHW fp32 CPU result, stats."""
    n, c0, h, wd = x.shape
    if int(split) == 2:
        xs, x1s = to_f8_operands(cuda, x), (to_f8_operands(cuda, x1) if x1 is not None else None)
    else:
        xs = K.nchw_to_nhwc_split(x.to(cuda), split=split)
        x1s = K.nchw_to_nhwc_split(x1.to(cuda), split=split) if x1 is not None else None
    ws = K.pack_conv_weight(w.to(cuda), transposed=transposed, split=split)
    cout = w.shape[1] if transposed else w.shape[0]
    kh, kw = w.shape[2:]
    d = K.make_conv_desc(n, h, wd, c0, cout, kh, kw, stride=stride, pad=pad, dil=dil,
                         cin1=0 if x1 is None else x1.shape[1], transposed=transposed, split=split, n_tile=n_tile,
                         halo=halo, pad_w=pad_w)
    out = torch.full((n, d.h_out, d.w_out, cout), float("nan"), dtype=torch.float32, device=cuda)
    st = torch.zeros((n, cout, 2), dtype=torch.float64, device=cuda) if stats else None
    plan = K.ConvPlan(d, xs, x1s, ws, out, st)
    plan.run()
    torch.cuda.synchronize()
    return K.nhwc_to_nchw(out).cpu(), (st.cpu() if stats else None)


def to_f8_operands(cuda, x):
    """NCHW fp32 -> (hi fp16 NHWC, fp8 pair blocks) through the norm kernel used as a plain converter."""
    raw = x.permute(0, 2, 3, 1).contiguous().to(cuda)
    hi = torch.empty(raw.shape, dtype=torch.float16, device=cuda)
    lo = torch.empty_like(hi)
    K.norm_act_nhwc(raw, None, None, None, False, None, y_hi=hi, y_lo=lo, lo_format=1)
    return hi, lo


def q8(t):
    return t.clamp(-448, 448).to(torch.float8_e4m3fn).double()


def emulate_f8(x, w, conv):
    """The arithmetic the fp16f8 mode is meant to perform, in float64 on the CPU:
    per-layer weight scale 2^E with max|w| * 2^E in [2^14, 2^15); x8 = e4m3(x / 16), xlo8 = e4m3(x_lo * 2^10),
    wlo8 = e4m3(w_lo * 2^(E+4)), w8 = e4m3(w * 2^(E-10))."""
    E = K.weight_exponent(w.abs().max())
    xh = x.half().double()
    wh = w.half().double()
    xl, wl = x.double() - xh, w.double() - wh
    main = conv(xh, wh)
    t2 = conv(q8(x / 16), q8((wl * 2.0 ** (E + 4)).float())) / 2.0 ** E
    t3 = conv(q8((xl * 1024).float()), q8(w * 2.0 ** (E - 10))) / 2.0 ** E
    return (main + t2 + t3).float()


def check_stats(st, ref):
    s = ref.double().sum(dim=(2, 3))
    q = (ref.double() ** 2).sum(dim=(2, 3))
    e1 = ((st[..., 0] - s).abs().max() / (s.abs().max() + 1e-9)).item()
    e2 = ((st[..., 1] - q).abs().max() / (q.abs().max() + 1e-9)).item()
    print("stats rel err: sum %.3e sumsq %.3e" % (e1, e2))
    assert e1 < 1e-3 and e2 < 1e-3


CASES = [
    # name, n, cin, cout, h, w, k, stride, pad, n_tile
    ("3x3_64_64_16x8", 1, 64, 64, 16, 8, 3, 1, 1, 0),
    ("3x3_64_64_32", 2, 64, 64, 32, 32, 3, 1, 1, 0),
    ("3x3_128_128_32", 2, 128, 128, 32, 32, 3, 1, 1, 0),
    ("3x3_512_512_32", 2, 512, 512, 32, 32, 3, 1, 1, 0),
    ("3x3_64_16_n16", 1, 64, 16, 32, 32, 3, 1, 1, 16),
    ("1x1_64_64", 1, 64, 64, 32, 32, 1, 1, 0, 0),
    ("3x3_s2_64_128_64", 2, 64, 128, 64, 64, 3, 2, 1, 0),
    ("3x3_s2_256_512_64", 1, 256, 512, 64, 64, 3, 2, 1, 0),
    ("3x3_ragged_40x24", 1, 64, 64, 40, 24, 3, 1, 1, 0),
    ("7x7_64_64", 1, 64, 64, 32, 32, 7, 1, 3, 0),
]


@pytest.mark.parametrize("split", [True, False])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv2d(cuda, case, split):
    name, n, cin, cout, h, w, k, stride, pad, n_tile = case
    x = rnd(n, cin, h, w, seed=1)
    wt = rnd(cout, cin, k, k, seed=2, scale=0.05)
    ref = F.conv2d(x, wt, stride=stride, padding=pad)
    got, st = run_conv(cuda, x, wt, stride=stride, pad=pad, split=split, n_tile=n_tile)
    rel = report(name + ("/split" if split else "/fast"), got, ref)
    assert rel < (2e-4 if split else 2e-2)
    if split:
        check_stats(st, ref)


HALO_CASES = [c for c in CASES if c[7] == 1 and c[6] in (3, 7) and c[8] == c[6] // 2] + [
    ("7x7_64_16_heads", 2, 64, 16, 48, 40, 7, 1, 3, 16),
    ("5x5_64_64", 1, 64, 64, 32, 32, 5, 1, 2, 0),
    ("3x3_128_64_64x64", 2, 128, 64, 64, 64, 3, 1, 1, 0),
]


@pytest.mark.parametrize("split", [True, False])
@pytest.mark.parametrize("case", HALO_CASES, ids=[c[0] for c in HALO_CASES])
def test_conv2d_halo(cuda, case, split):
    """Halo variant: one activation tile (+halo) per chunk, every tap a shifted UMMA descriptor window."""
    name, n, cin, cout, h, w, k, stride, pad, n_tile = case
    x = rnd(n, cin, h, w, seed=31)
    wt = rnd(cout, cin, k, k, seed=32, scale=0.05)
    ref = F.conv2d(x, wt, stride=1, padding=pad)
    got, st = run_conv(cuda, x, wt, stride=1, pad=pad, split=split, n_tile=n_tile, halo=True)
    rel = report(name + "/halo" + ("/split" if split else "/fast"), got, ref)
    assert rel < (2e-4 if split else 2e-2)
    if split:
        check_stats(st, ref)


@pytest.mark.parametrize("split", [True, False])
def test_conv_concat_inputs_halo(cuda, split):
    a, b = rnd(2, 64, 32, 32, seed=5), rnd(2, 128, 32, 32, seed=6)
    wt = rnd(64, 192, 3, 3, seed=7, scale=0.05)
    ref = F.conv2d(torch.cat([a, b], dim=1), wt, padding=1)
    got, _ = run_conv(cuda, a, wt, split=split, x1=b, halo=True)
    assert report("concat/halo", got, ref) < (2e-4 if split else 2e-2)


@pytest.mark.parametrize("split", [True, False])
def test_conv_transpose(cuda, split):
    for cin, cout, h in ((128, 64, 32), (512, 256, 32)):
        x = rnd(2, cin, h, h, seed=3)
        wt = rnd(cin, cout, 3, 3, seed=4, scale=0.05)
        ref = F.conv_transpose2d(x, wt, stride=2, padding=1, output_padding=1)
        got, st = run_conv(cuda, x, wt, stride=2, pad=1, transposed=True, split=split)
        rel = report("convT_%d_%d" % (cin, cout), got, ref)
        assert rel < (2e-4 if split else 2e-2)
        if split:
            check_stats(st, ref)


@pytest.mark.parametrize("split", [True, False])
def test_conv_concat_inputs(cuda, split):
    """skippers: conv(cat[skip, d]) without materialising the cat (networks/generator.py:177-179)."""
    a, b = rnd(2, 64, 32, 32, seed=5), rnd(2, 128, 32, 32, seed=6)
    wt = rnd(64, 192, 3, 3, seed=7, scale=0.05)
    ref = F.conv2d(torch.cat([a, b], dim=1), wt, padding=1)
    got, _ = run_conv(cuda, a, wt, split=split, x1=b)
    assert report("concat", got, ref) < (2e-4 if split else 2e-2)


@pytest.mark.parametrize("halo", [False, True])
@pytest.mark.parametrize("split", [True, False])
@pytest.mark.parametrize("size", [32, 256])
def test_stem_7x7_rowk(cuda, split, size, halo):
    """7x7 stem (6 -> 64 channels) through the row-K layout (networks/generator.py:80-84)."""
    n = 2
    x = rnd(n, 6, size, size, seed=8)
    wt = rnd(64, 6, 7, 7, seed=9, scale=0.05)
    ref = F.conv2d(x, wt, padding=3)
    pitch = size + 8
    xs = K.nchw_to_nhwc_split(x.to(cuda), c_pad=8, pad_hw=(3, 3, 3, 5), split=split)
    ws = K.pack_conv_weight_rowk(wt.to(cuda), split=split)
    d = K.make_conv_desc(n, size, size, 8, 64, 7, 7, stride=1, pad=3, split=split, rowk=True, row_pitch=pitch, halo=halo)
    out = torch.full((n, size, size, 64), float("nan"), dtype=torch.float32, device=cuda)
    st = torch.zeros((n, 64, 2), dtype=torch.float64, device=cuda)
    K.ConvPlan(d, xs, None, ws, out, st).run()
    torch.cuda.synchronize()
    got = K.nhwc_to_nchw(out).cpu()
    assert report("stem_rowk_%d" % size, got, ref) < (2e-4 if split else 2e-2)
    check_stats(st.cpu(), ref)


def test_heads_7x7(cuda):
    x = rnd(2, 64, 64, 96, seed=10)
    w_img, w_att = rnd(3, 64, 7, 7, seed=11, scale=0.02), rnd(1, 64, 7, 7, seed=12, scale=0.02)
    ref = F.conv2d(x, torch.cat([w_img, w_att]), padding=3)
    xn = x.permute(0, 2, 3, 1).contiguous().to(cuda)
    raw = K.conv7x7_heads_nhwc(xn, K.pack_head_weights(w_img.to(cuda), w_att.to(cuda)))
    got = raw.permute(0, 3, 1, 2).cpu()
    assert report("heads", got, ref) < 1e-5
    bg = rnd(1, 3, 64, 96, seed=13)
    color, mask, pred = K.heads_composite(raw, bg.to(cuda))
    rc, rm = torch.tanh(ref[:, :3]), torch.sigmoid(ref[:, 3:])
    assert (color.cpu() - rc).abs().max() < 1e-5 and (mask.cpu() - rm).abs().max() < 1e-5
    assert (pred.cpu() - (rm * bg + (1 - rm) * rc)).abs().max() < 1e-5       # models/imitator.py:331


def test_norm_act_warp(cuda):
    """IN + ReLU + residual + LWB warp-add (networks/generator.py:13-20,283-295,303-320)."""
    n, c, h, w = 3, 64, 32, 32
    raw = rnd(n, c, h, w, seed=14) * 2 + 0.5
    gamma, beta = 1 + 0.1 * rnd(c, seed=15), 0.1 * rnd(c, seed=16)
    res = rnd(n, c, h, w, seed=17)
    src = rnd(1, c, h, w, seed=18)
    from impersonator_b200 import synthetic as S
    T = S.synthetic_flow(n, 256, seed=3)
    for ac in (False, True):
        Ts = F.interpolate(T.permute(0, 3, 1, 2), size=(h, w), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
        warp = F.grid_sample(src.expand(n, -1, -1, -1), Ts, mode="bilinear", padding_mode="zeros", align_corners=ac)
        ref = F.relu(F.instance_norm(raw, weight=gamma, bias=beta, eps=1e-5)) + res + warp
        nh = lambda t: t.permute(0, 2, 3, 1).contiguous().to(cuda)
        raw_d = nh(raw)
        st = K.instance_stats_nhwc(raw_d)
        y = torch.empty_like(raw_d)
        hi = torch.empty(raw_d.shape, dtype=torch.float16, device=cuda)
        lo = torch.empty_like(hi)
        ws = torch.empty((n, c, 2), dtype=torch.float32, device=cuda)
        K.norm_act_nhwc(raw_d, st, gamma.to(cuda), beta.to(cuda), True, ws, residual=nh(res), warp_src=nh(src),
                        T=T.to(cuda), align_corners=ac, y_f32=y, y_hi=hi, y_lo=lo)
        got = y.permute(0, 3, 1, 2).cpu()
        assert report("norm_act ac=%s" % ac, got, ref) < 2e-5
        rec = (hi.float() + lo.float()).permute(0, 3, 1, 2).cpu()
        assert (rec - got).abs().max() < 1e-5


@pytest.mark.parametrize("ac", [False, True])
def test_warp_nchw(cuda, ac):
    from impersonator_b200 import synthetic as S
    x = rnd(1, 24, 64, 64, seed=19)
    T = S.synthetic_flow(2, 256, seed=4)
    Ts = F.interpolate(T.permute(0, 3, 1, 2), size=(64, 64), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    ref = F.grid_sample(x.expand(2, -1, -1, -1), Ts, mode="bilinear", padding_mode="zeros", align_corners=ac)
    got = K.warp_nchw(x.to(cuda), T.to(cuda), align_corners=ac).cpu()
    assert report("transform", got, ref) < 2e-5
    img = rnd(2, 3, 256, 256, seed=20)
    ref2 = F.grid_sample(img, T, mode="bilinear", padding_mode="zeros", align_corners=ac)
    got2 = K.warp_nchw(img.to(cuda), T.to(cuda), align_corners=ac).cpu()
    assert report("stn", got2, ref2) < 2e-5


def test_direct_conv(cuda):
    x = rnd(1, 5, 40, 40, seed=21)
    for (co, k, s, p, d) in ((8, 5, 1, 2, 1), (6, 4, 2, 1, 1), (7, 3, 1, 4, 4), (3, 1, 1, 0, 1)):
        wt, b = rnd(co, 5, k, k, seed=22, scale=0.1), rnd(co, seed=23)
        ref = F.conv2d(x, wt, b, stride=s, padding=p, dilation=d)
        got = K.conv2d_direct_nchw(x.to(cuda), wt.to(cuda), b.to(cuda), stride=s, pad=p, dil=d).cpu()
        assert report("direct k%d s%d d%d" % (k, s, d), got, ref) < 1e-5


F8_CASES = [c for c in CASES if c[2] % 64 == 0 and c[9] != 16]


@pytest.mark.parametrize("case", F8_CASES, ids=[c[0] for c in F8_CASES])
def test_conv2d_fp16f8(cuda, case):
    """split = 2: x_hi*w_hi on the fp16 path + (x*w_lo, x_lo*w) as e4m3 MMAs into the same accumulator."""
    name, n, cin, cout, h, w, k, stride, pad, n_tile = case
    x = rnd(n, cin, h, w, seed=1)
    wt = rnd(cout, cin, k, k, seed=2, scale=0.05)
    ref = F.conv2d(x, wt, stride=stride, padding=pad)
    got, st = run_conv(cuda, x, wt, stride=stride, pad=pad, split=2, n_tile=n_tile)
    rel = report(name + "/fp16f8 vs fp32", got, ref)
    assert rel < 3e-4
    emu = emulate_f8(x, wt, lambda a, b: F.conv2d(a, b, stride=stride, padding=pad))
    rel_e = report(name + "/fp16f8 vs its float64 emulation", got, emu)
    assert rel_e < 3e-5                      # only the fp32 accumulation order of the 2^15-scaled sums differs
    check_stats(st, ref)


def test_conv_transposed_and_concat_fp16f8(cuda):
    x = rnd(2, 128, 32, 32, seed=5)
    wt = rnd(128, 64, 3, 3, seed=6, scale=0.05)
    ref = F.conv_transpose2d(x, wt, stride=2, padding=1, output_padding=1)
    got, _ = run_conv(cuda, x, wt, stride=2, pad=1, transposed=True, split=2, stats=False)
    assert report("convT 128->64 /fp16f8", got, ref) < 3e-4
    emu = emulate_f8(x, wt, lambda a, b: F.conv_transpose2d(a, b, stride=2, padding=1, output_padding=1))
    assert report("convT 128->64 /fp16f8 vs emulation", got, emu) < 1e-5
    a, b = rnd(1, 64, 64, 64, seed=7), rnd(1, 64, 64, 64, seed=8)
    w2 = rnd(64, 128, 3, 3, seed=9, scale=0.05)
    ref = F.conv2d(torch.cat([a, b], dim=1), w2, padding=1)
    got, st = run_conv(cuda, a, w2, split=2, x1=b)
    assert report("concat 64+64->64 /fp16f8", got, ref) < 3e-4
    check_stats(st, ref)


@pytest.mark.parametrize("split", [1, 2])
def test_conv_homogeneity_and_batch_invariance_at_full_batch(cuda, split):
    """Size-independent properties at the BASELINE batch (16 x 512 x 32 x 32, the residual-block layer): conv(2x) = 2 conv(x)
    up to the fp16-subnormal rounding of the lo operands (|lo| < 6e-5 loses bits, so not bit-exact), and every image of a
    batch of identical images gets the same bits (same tile schedule, same accumulation order)."""
    x1 = rnd(1, 512, 32, 32, seed=3)
    x = x1.expand(16, -1, -1, -1).contiguous()
    wt = rnd(512, 512, 3, 3, seed=4, scale=0.02)
    y, _ = run_conv(cuda, x, wt, split=split, stats=False)
    y2, _ = run_conv(cuda, 2 * x, wt, split=split, stats=False)
    assert report("conv(2x) vs 2 conv(x)", y2, 2 * y) < 2e-5
    assert torch.equal(y[0], y[15]) and torch.equal(y[0], y[7])
    ref = F.conv2d(x1, wt, padding=1)
    assert report("512->512 @32 batch 16 (image 0)", y[:1], ref) < (2e-4 if split == 1 else 3e-4)


# Shapes the round-2 callers add: tiny images under one 16x8 tile (HMR layer3/4: 14x14, 7x7), 1x1 filters with up to 2048
# input channels, N tile 32 (folded heads, gated 16-channel layers, q/k/v), dilation 16, 5x5, 4x4 stride 2.
R2_CASES = [
    # name, n, cin, cout, h, w, k, stride, pad, dil, n_tile
    ("hmr_1x1_2048_512_7x7", 3, 2048, 512, 7, 7, 1, 1, 0, 1, 0),
    ("hmr_3x3_512_512_7x7", 3, 512, 512, 7, 7, 3, 1, 1, 1, 0),
    ("hmr_3x3_s2_256_256_14", 3, 256, 256, 14, 14, 3, 2, 1, 1, 0),
    ("hmr_1x1_64_256_56", 1, 64, 256, 56, 56, 1, 1, 0, 1, 0),
    ("hmr_3x3_s2_64_64_56", 2, 64, 64, 56, 56, 3, 2, 1, 1, 0),
    ("n32_3x3_64_32", 1, 64, 32, 40, 24, 3, 1, 1, 1, 0),
    ("attn_1x1_128_160", 1, 128, 160, 64, 64, 1, 1, 0, 1, 0),
    ("inp_3x3_dil16_128_256", 1, 128, 256, 64, 64, 3, 1, 16, 16, 0),
    ("inp_5x5_64_64", 1, 64, 64, 64, 64, 5, 1, 2, 1, 0),
    ("inp_4x4_s2_64_128", 1, 64, 128, 64, 64, 4, 2, 1, 1, 0),
]


@pytest.mark.parametrize("split", [1, 2])
@pytest.mark.parametrize("case", R2_CASES, ids=[c[0] for c in R2_CASES])
def test_conv2d_round2_shapes(cuda, case, split):
    name, n, cin, cout, h, w, k, stride, pad, dil, n_tile = case
    x = rnd(n, cin, h, w, seed=11)
    wt = rnd(cout, cin, k, k, seed=12, scale=0.05)
    ref = F.conv2d(x, wt, stride=stride, padding=pad, dilation=dil)
    got, st = run_conv(cuda, x, wt, stride=stride, pad=pad, dil=dil, split=split, n_tile=n_tile)
    rel = report(name + "/split%d" % split, got, ref)
    assert rel < 3e-4
    check_stats(st, ref)


@pytest.mark.parametrize("split", [1, 2])
def test_conv_7x1_folded_heads(cuda, split):
    """The 7x7 heads as a 7x1 filter with N = 7 columns x 4 channels (generator.fold_head_weights) + the column sum in
    lwb_heads_composite(folded_kw=7), against F.conv2d + tanh / sigmoid."""
    from impersonator_b200.generator import fold_head_weights
    n, h, w = 2, 48, 40
    x = rnd(n, 64, h, w, seed=21)
    w_img, w_att = rnd(3, 64, 7, 7, seed=22, scale=0.02), rnd(1, 64, 7, 7, seed=23, scale=0.02)
    folded = fold_head_weights(w_img, w_att)
    raw, _ = run_conv(cuda, x, folded, stride=1, pad=3, pad_w=0, split=split, n_tile=32, stats=False)
    raw_nhwc = raw.permute(0, 2, 3, 1).contiguous().to(cuda)
    color, mask, _ = K.heads_composite(raw_nhwc, None, folded_kw=7)
    ref_c = torch.tanh(F.conv2d(x, w_img, padding=3))
    ref_m = torch.sigmoid(F.conv2d(x, w_att, padding=3))
    d = max((color.cpu() - ref_c).abs().max().item(), (mask.cpu() - ref_m).abs().max().item())
    print("folded 7x1 heads split %d vs torch: %.3e" % (split, d))
    assert d < 2e-4


@pytest.mark.parametrize("variant", ["plain", "res_warp"])
@pytest.mark.parametrize("split", [1, 2])
def test_conv_fused_instance_norm(cuda, split, variant, monkeypatch):
    """lwb_conv_plan_fuse_norm against conv + lwb_norm_act_nhwc on the same operands: 512 -> 512 @32x32, batch 3 (24 tiles per
    N tile: several CTAs wait on every unit), twice in a row (counters / statistics re-zeroed)."""
    monkeypatch.setenv("LWB_YHALO", "0")                  # the fused epilogue lives in the non-halo 2-CTA kernel
    n, c, h, w = 3, 512, 32, 32
    x = rnd(n, c, h, w, seed=41)
    wt = rnd(c, c, 3, 3, seed=42, scale=0.03)
    gamma, beta = (1 + 0.1 * rnd(c, seed=43)).to(cuda), (0.1 * rnd(c, seed=44)).to(cuda)
    if int(split) == 2:
        xs = to_f8_operands(cuda, x)
    else:
        xs = K.nchw_to_nhwc_split(x.to(cuda), split=True)
    ws = K.pack_conv_weight(wt.to(cuda), split=split)
    d = K.make_conv_desc(n, h, w, c, c, 3, 3, stride=1, pad=1, split=split)
    lo_format = 1 if int(split) == 2 else 0
    res = rnd(n, h, w, c, seed=45).to(cuda) if variant == "res_warp" else None
    src = rnd(1, h, w, c, seed=46).to(cuda) if variant == "res_warp" else None
    T = (torch.rand(n, 64, 64, 2, generator=torch.Generator().manual_seed(47)) * 2.4 - 1.2).to(cuda) if variant == "res_warp" else None

    def outs():
        return (torch.empty(n, h, w, c, device=cuda), torch.empty(n, h, w, c, dtype=torch.float16, device=cuda),
                torch.empty(n, h, w, c, dtype=torch.float16, device=cuda))
    # separate pass
    raw = torch.empty(n, h, w, c, device=cuda)
    st = torch.zeros(n, c, 2, dtype=torch.float64, device=cuda)
    K.ConvPlan(d, xs, None, ws, raw, st).run()
    y0, hi0, lo0 = outs()
    wsb = torch.empty(n, c, 2, device=cuda)
    K.norm_act_nhwc(raw, st, gamma, beta, variant == "plain", wsb, residual=res, warp_src=src, T=T, align_corners=True,
                    y_f32=y0, y_hi=hi0, y_lo=lo0, lo_format=lo_format)
    # fused
    st1 = torch.zeros(n, c, 2, dtype=torch.float64, device=cuda)
    ctr = torch.zeros(n * 4, dtype=torch.int32, device=cuda)
    raw1 = torch.full((n, h, w, c), float("nan"), device=cuda)
    plan = K.ConvPlan(d, xs, None, ws, raw1, st1)
    y1, hi1, lo1 = outs()
    ok = plan.fuse_norm(gamma, beta, variant == "plain", ctr, residual=res, warp_src=src, T=T, align_corners=True,
                        y_f32=y1, y_hi=hi1, y_lo=lo1, lo_format=lo_format)
    assert ok, "a 32x32 split-mode plan must be fusable"
    for _ in range(2):
        st1.zero_()
        ctr.zero_()
        plan.run()
    torch.cuda.synchronize()
    dy = (y1 - y0).abs().max().item()
    print("fused norm %s split %d: y max-abs diff %.3e (scale %.2f)" % (variant, split, dy, y0.abs().max().item()))
    assert dy < 1e-5
    assert (hi1.float() - hi0.float()).abs().max().item() < 2e-3
    assert float(lo1.view(torch.uint8).ne(lo0.view(torch.uint8)).float().mean()) < 0.01
    assert int(ctr[:n * 2].max()) == 8 and int(ctr[:n * 2].min()) == 8          # every unit saw its 8 tiles


@pytest.mark.parametrize("split", [1, 2])
@pytest.mark.parametrize("cin,cout,h", [(128, 64, 32), (256, 128, 24)])
def test_conv_transposed_merged_phases(cuda, split, cin, cout, h):
    """lwb_conv_desc.transposed = 2: ConvTranspose2d(k3, s2, p1, op1) as ONE stride-1 pass with the four sub-pixel phases
    stacked on N (generator.merge_transposed_weight), against F.conv_transpose2d; statistics shared by the four phases."""
    from impersonator_b200.generator import merge_transposed_weight
    n, w = 2, 40
    x = rnd(n, cin, h, w, seed=51)
    wt = rnd(cin, cout, 3, 3, seed=52, scale=0.05)
    ref = F.conv_transpose2d(x, wt, stride=2, padding=1, output_padding=1)
    xs = to_f8_operands(cuda, x) if split == 2 else K.nchw_to_nhwc_split(x.to(cuda), split=True)
    ws = K.pack_conv_weight(merge_transposed_weight(wt.to(cuda)), split=split)
    d = K.make_conv_desc(n, h, w, cin, cout, 3, 3, stride=2, pad=1, transposed=True, split=split)
    d.transposed = 2
    out = torch.full((n, 2 * h, 2 * w, cout), float("nan"), device=cuda)
    st = torch.zeros(n, cout, 2, dtype=torch.float64, device=cuda)
    plan = K.ConvPlan(d, xs, None, ws, out, st)
    assert plan.num_launches == 1
    plan.run()
    torch.cuda.synchronize()
    got = K.nhwc_to_nchw(out).cpu()
    assert report("merged convT %d->%d split %d" % (cin, cout, split), got, ref) < 3e-4
    check_stats(st.cpu(), ref)
