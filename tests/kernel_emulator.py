"""TEST INFRASTRUCTURE ONLY -- torch-CPU stand-ins for the C-ABI front-ends of impersonator_b200.kernels.

The host mirrors (generator / hmr / inpaintor streams) are pure orchestration: they bind buffers, fold BatchNorms, stack
gated filters, fold the 7x7 heads into a 7x1 filter, chain conv plans and epilogues.  None of that needs a GPU to be
wrong.  ``install(monkeypatch)`` replaces every kernel front-end the streams call by a plain torch implementation of the
SAME contract (include/lwb_b200.h), so the CPU suite can run the whole host logic against the oracles.  The emulation
works on the fp16 hi/lo operand pairs (run it with LWB_PRECISION=fp16x3); it is never used by the product.
"""
import torch
import torch.nn.functional as F

from impersonator_b200 import kernels as K


def _pair_to_f32(pair):
    hi, lo = pair
    return hi.float() + (lo.float() if lo is not None else 0.0)


def _emit(y, y_f32, y_hi, y_lo):
    if y_f32 is not None:
        y_f32.copy_(y)
    if y_hi is not None:
        c = y.shape[-1]
        y_hi.zero_()
        y_hi[..., :c] = y.half()
        if y_lo is not None:
            y_lo.zero_()
            y_lo[..., :c] = (y - y_hi[..., :c].float()).half()


class PackedWeight(tuple):
    pass


def pack_conv_weight(w, transposed=False, cout_pad=None, cin_pad=None, split=True, absmax=None):
    w = w.detach().float()
    if transposed:
        cin, cout = w.shape[:2]
    else:
        cout, cin = w.shape[:2]
    cout_pad, cin_pad = cout_pad or cout, cin_pad or cin
    full = torch.zeros((cin_pad, cout_pad) + tuple(w.shape[2:])) if transposed else torch.zeros((cout_pad, cin_pad) + tuple(w.shape[2:]))
    if transposed:
        full[:cin, :cout] = w
    else:
        full[:cout, :cin] = w
    pw = PackedWeight((full, transposed))
    pw.w_exp = 15
    return pw


def pack_conv_weight_rowk(w, cout_pad=None, cpx=8, kxs=8, split=True):
    pw = PackedWeight((w.detach().float(), "rowk"))
    pw.w_exp = 15
    return pw


class ConvPlan(object):
    """lwb_conv_plan_create / run on NHWC hi/lo pairs -> raw fp32 NHWC (+ InstanceNorm sums in f64)."""

    def __init__(self, desc, x0, x1, w, out_raw, stats):
        self.desc, self.x0, self.x1, self.w, self.out, self.stats = desc, x0, x1, w, out_raw, stats
        self.num_launches = 1
        self.flops = 0.0
        self.label = "emulated"
        self.prof_class = "conv"

    def fuse_norm(self, *a, **k):
        return False

    def run(self):
        d = self.desc
        x = _pair_to_f32(self.x0)
        if self.x1 is not None:
            x = torch.cat([x, _pair_to_f32(self.x1)], dim=-1)
        w, kind = self.w
        if kind == "rowk":
            # padded NHWC8 input [n, h+6, w+8, 8] holding pixel (y,x) at (y+3, x+3); 7x7 stride 1
            x = x[:, :, :d.w_in + 6, :w.shape[1]].permute(0, 3, 1, 2)
            y = F.conv2d(x, w)
        else:
            x = x.permute(0, 3, 1, 2)
            if d.transposed == 2:
                # merged transposed conv (lwb_conv_desc.transposed = 2): weights [4*cout, cin, 2, 2], tap (dy, dx) reads
                # in[y+dy, x+dx] (zero beyond the border), column block 2a+b is output pixel (2y+a, 2x+b)
                y4 = F.conv2d(F.pad(x, (0, 1, 0, 1)), w)
                n_, c4, h_, w_ = y4.shape
                y = y4.view(n_, 2, 2, c4 // 4, h_, w_).permute(0, 3, 4, 1, 5, 2).reshape(n_, c4 // 4, 2 * h_, 2 * w_)
            elif kind:
                y = F.conv_transpose2d(x, w, stride=2, padding=1, output_padding=1)
            else:
                pw = d.pad_w if d.pad_w >= 0 else d.pad
                y = F.conv2d(x, w, stride=d.stride, padding=(d.pad, pw), dilation=d.dil)
        y = y.permute(0, 2, 3, 1)
        assert tuple(y.shape) == tuple(self.out.shape), (tuple(y.shape), tuple(self.out.shape))
        self.out.copy_(y)
        if self.stats is not None:
            self.stats[..., 0] += y.double().sum(dim=(1, 2))
            self.stats[..., 1] += (y.double() ** 2).sum(dim=(1, 2))


def _warp(src, T, h, w, align_corners):
    Ts = F.interpolate(T.permute(0, 3, 1, 2), size=(h, w), mode='bilinear', align_corners=True).permute(0, 2, 3, 1)
    x = src.permute(0, 3, 1, 2)
    if x.shape[0] != T.shape[0]:
        x = x.expand(T.shape[0], -1, -1, -1)
    return F.grid_sample(x, Ts, mode='bilinear', padding_mode='zeros', align_corners=bool(align_corners)).permute(0, 2, 3, 1)


def norm_act_nhwc(raw, stats, gamma, beta, relu, ws, eps=1e-5, residual=None, warp_src=None, T=None, align_corners=False,
                  y_f32=None, y_hi=None, y_lo=None, lo_format=0, post_scale=None, post_shift=None, post_relu=False,
                  res_step=1, range_flag=None):
    n, h, w, c = raw.shape
    v = raw.double()
    if stats is not None:
        mean = stats[..., 0] / (h * w)
        var = (stats[..., 1] / (h * w) - mean * mean).clamp(min=0)
        scale = (gamma.double() if gamma is not None else 1.0) / torch.sqrt(var + eps)
        shift = (beta.double() if beta is not None else 0.0) - mean * scale
        v = v * scale[:, None, None, :] + shift[:, None, None, :]
    elif gamma is not None or beta is not None:
        v = v * (gamma.double() if gamma is not None else 1.0) + (beta.double() if beta is not None else 0.0)
    v = v.float()
    if relu:
        v = F.relu(v)
    if residual is not None:
        v = v + residual[:, ::res_step, ::res_step, :]
    if warp_src is not None:
        v = v + _warp(warp_src, T, h, w, align_corners)
    ops = v
    if post_scale is not None:
        ops = v * post_scale + post_shift
        if post_relu:
            ops = F.relu(ops)
    if y_f32 is not None:
        y_f32.copy_(v)
    _emit(ops, None, y_hi, y_lo)
    if range_flag is not None and y_hi is not None:
        m = float(ops.abs().max())
        if m >= 1024:
            range_flag |= 3 if m >= 60000 else 1


def nchw_to_nhwc_split(x, c_pad=None, pad_hw=(0, 0, 0, 0), hi=None, lo=None, split=True):
    n, c, h, w = x.shape
    c_pad = c_pad or c
    top, bottom, left, right = pad_hw
    full = torch.zeros((n, h + top + bottom, w + left + right, c_pad))
    full[:, top:top + h, left:left + w, :c] = x.permute(0, 2, 3, 1)
    if hi is None:
        hi = torch.empty(full.shape, dtype=torch.float16)
        lo = torch.empty_like(hi) if split else None
    hi.copy_(full.half())
    if lo is not None:
        lo.copy_((full - hi.float()).half())
    return hi, lo


def nhwc_to_nchw(x, c=None, out=None):
    c = c or x.shape[3]
    y = x[..., :c].permute(0, 3, 1, 2).contiguous()
    if out is not None:
        out.copy_(y)
        return out
    return y


def conv2d_direct_nchw(x, w, bias=None, stride=1, pad=0, dil=1):
    return F.conv2d(x, w, bias, stride=stride, padding=pad, dilation=dil)


def gated_bn_nchw(ab, act, scale=None, shift=None):
    c = ab.shape[1] // 2
    a, g = ab[:, :c], ab[:, c:]
    a = F.leaky_relu(a, 0.2) if act == 2 else (F.relu(a) if act == 1 else a)
    y = a * torch.sigmoid(g)
    if scale is not None:
        y = y * scale[None, :, None, None] + shift[None, :, None, None]
    return y


def gated_act_nhwc(raw, c, bias, act, scale, shift, upsample=1, clamp=False, y_f32=None, y_hi=None, y_lo=None,
                   lo_format=0, range_flag=None):
    a, g = raw[..., :c], raw[..., c:2 * c]
    if bias is not None:
        a, g = a + bias[:c], g + bias[c:]
    a = F.leaky_relu(a, 0.2) if act == 2 else a
    y = a * torch.sigmoid(g)
    if scale is not None:
        y = y * scale + shift
    if clamp:
        y = y.clamp(-1, 1)
    if upsample == 2:
        y = y.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    _emit(y, y_f32, y_hi, y_lo)


def self_attention_nhwc(qkv, bias, x, gamma, dq=16, out=None):
    n, h, w, ld = qkv.shape
    dv = x.shape[3]
    t = (qkv[..., :2 * dq + dv] + bias).view(n, h * w, 2 * dq + dv)
    q, k, v = t[..., :dq], t[..., dq:2 * dq], t[..., 2 * dq:]
    att = torch.softmax(torch.bmm(q, k.transpose(1, 2)), dim=-1)
    y = (gamma * torch.bmm(att, v) + x.view(n, h * w, dv)).view(n, h, w, dv)
    if out is not None:
        out.copy_(y)
        return out
    return y


def maxpool_nchw_to_nhwc(x, k, stride, out=None):
    y = F.max_pool2d(x, kernel_size=k, stride=stride, ceil_mode=True).permute(0, 2, 3, 1).contiguous()
    if out is not None:
        out.copy_(y)
        return out
    return y


def global_avgpool_nhwc(x, scale=None, shift=None, relu=False, out=None, ld_out=None):
    v = x
    if scale is not None:
        v = v * scale + shift
    if relu:
        v = F.relu(v)
    y = v.mean(dim=(1, 2))
    if out is None:
        return y
    out[:, :y.shape[1]] = y
    return out


def linear(x, w, bias=None, relu=False, out=None, accumulate=False):
    y = F.linear(x, w, bias)
    if relu:
        y = F.relu(y)
    if out is None:
        return y
    if accumulate:
        out += y
    else:
        out.copy_(y)
    return out


def pack_head_weights(w_img, w_att):
    return torch.cat([w_img, w_att], dim=0).float()


def conv7x7_heads_nhwc(x, w4, out=None):
    y = F.conv2d(x.permute(0, 3, 1, 2), w4, padding=3).permute(0, 2, 3, 1)
    if out is not None:
        out.copy_(y)
        return out
    return y


def heads_composite(raw, bg=None, want_color=True, want_mask=True, color=None, mask=None, pred=None, want_pred=True,
                    pred_hwc=None, pred_u8=None, folded_kw=0, range_flag=None):
    n, h, w, cs = raw.shape
    if folded_kw:
        r = torch.zeros(n, h, w, 4)
        for kx in range(folded_kw):
            sh = kx - folded_kw // 2                       # out[y, x] += raw[y, x + sh, kx*4 : kx*4+4]
            src = raw[..., kx * 4:kx * 4 + 4]
            if sh >= 0:
                r[:, :, :w - sh] += src[:, :, sh:]
            else:
                r[:, :, -sh:] += src[:, :, :w + sh]
    else:
        r = raw[..., :4]
    col = torch.tanh(r[..., :3]).permute(0, 3, 1, 2)
    m = torch.sigmoid(r[..., 3:4]).permute(0, 3, 1, 2)
    p = None
    if bg is not None:
        p = m * bg + (1 - m) * col
    outs = []
    for given, val in ((color, col), (mask, m), (pred, p)):
        if given is not None and val is not None:
            given.copy_(val)
            outs.append(given)
        else:
            outs.append(val.contiguous() if val is not None else None)
    if pred_hwc is not None:
        pred_hwc.copy_(p.permute(0, 2, 3, 1))
    return tuple(outs)


def correspond(cam, verts, face_idx, image_size, map_fn, src_p2verts, src_img=None, align_corners=None,
               want_f2verts=False, near=None, far=None, out=None):
    """lwb_correspond's contract through the CPU oracle (oracle/nmr_ref.py + the C rasterizer restatement)."""
    from oracle import nmr_ref
    ac = K.default_align_corners() if align_corners is None else align_corners
    img = src_img if src_img is not None else torch.zeros(1, 3, image_size, image_size)
    c = nmr_ref.correspond(cam, verts, face_idx, map_fn, src_p2verts, img, image_size, align_corners=ac)
    res = dict(fim=c["fim"], wim=c["wim"], T=c["T"], tsf_inputs=c["tsf_inputs"].contiguous(),
               f2verts=c["f2verts"] if want_f2verts else None)
    res["tsf_img"] = res["tsf_inputs"][:, :3]
    res["cond"] = res["tsf_inputs"][:, 3:]
    return res


def warp_nchw(x, T, align_corners=None, out=None, accumulate=False):
    ac = K.default_align_corners() if align_corners is None else align_corners
    y = torch.nn.functional.grid_sample(x.expand(T.shape[0], -1, -1, -1), T, mode='bilinear', padding_mode='zeros',
                                        align_corners=ac)
    if out is not None:
        out.copy_(out + y if accumulate else y)
        return out
    return y


def install_tasks(monkeypatch):
    """install() + the correspondence / warp front-ends and the renderer's CUDA-only guard: enough to run the task classes'
    personalize / view / swap on CPU (Imitator.inference itself drives CUDA streams and stays GPU-only)."""
    from impersonator_b200 import nmr
    install(monkeypatch)
    monkeypatch.setattr(K, "correspond", correspond)
    monkeypatch.setattr(K, "warp_nchw", warp_nchw)

    def _correspond(self, cam, vertices, src_p2verts, src_img, want_f2verts=False, align_corners=None):
        if src_p2verts is None:
            src_p2verts = torch.zeros((1, self.nf, 3, 2), dtype=torch.float32)
        return correspond(cam.float(), vertices.float(), self.faces, self.image_size, self.map_fn, src_p2verts.float(),
                          src_img, align_corners=align_corners, want_f2verts=want_f2verts)
    monkeypatch.setattr(nmr.SMPLRenderer, "_correspond", _correspond)


def install(monkeypatch):
    for name in ("pack_conv_weight", "pack_conv_weight_rowk", "ConvPlan", "norm_act_nhwc", "nchw_to_nhwc_split", "nhwc_to_nchw",
                 "conv2d_direct_nchw", "gated_bn_nchw", "gated_act_nhwc", "self_attention_nhwc", "maxpool_nchw_to_nhwc",
                 "global_avgpool_nhwc", "linear", "pack_head_weights", "conv7x7_heads_nhwc", "heads_composite"):
        monkeypatch.setattr(K, name, globals()[name])
    monkeypatch.setenv("LWB_PRECISION", "fp16x3")
