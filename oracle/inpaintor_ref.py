"""TEST INFRASTRUCTURE ONLY -- functional torch-fp32 restatement of networks/inpaintor.py.

  GatedConv2dWithActivation   networks/inpaintor.py:12-47   (conv, mask conv, LeakyReLU(0.2)*sigmoid, eval BatchNorm)
  GatedDeConv2dWithActivation :50-68   (F.interpolate(scale_factor=2) = nearest, then gated conv)
  SelfAttention               :71-107
  InpaintSANet.forward        :178-202
Takes the reference's state_dict (same keys); validated against the imported reference module and
pinned by tests/golden/inpaintor.npz (tests/golden/make_inpaintor_golden.py).
"""
import torch
import torch.nn.functional as F

# (kernel, stride, dilation, deconv, activation) per layer, read off networks/inpaintor.py:117-176
COARSE = [(5, 1, 1, 0, 1), (4, 2, 1, 0, 1), (3, 1, 1, 0, 1), (4, 2, 1, 0, 1), (3, 1, 1, 0, 1), (3, 1, 1, 0, 1),
          (3, 1, 2, 0, 1), (3, 1, 4, 0, 1), (3, 1, 8, 0, 1), (3, 1, 16, 0, 1), (3, 1, 1, 0, 1), (3, 1, 1, 0, 1),
          (3, 1, 1, 1, 1), (3, 1, 1, 0, 1), (3, 1, 1, 1, 1), (3, 1, 1, 0, 1), (3, 1, 1, 0, 0)]
REFINE = [(5, 1, 1, 0, 1), (4, 2, 1, 0, 1), (3, 1, 1, 0, 1), (4, 2, 1, 0, 1), (3, 1, 1, 0, 1), (3, 1, 1, 0, 1),
          (3, 1, 1, 0, 1), (3, 1, 2, 0, 1), (3, 1, 4, 0, 1), (3, 1, 8, 0, 1), (3, 1, 16, 0, 1)]
UPSAMPLE = [(3, 1, 1, 0, 1), (3, 1, 1, 0, 1), (3, 1, 1, 1, 1), (3, 1, 1, 0, 1), (3, 1, 1, 1, 1), (3, 1, 1, 0, 1), (3, 1, 1, 0, 0)]


def _pad(k, s, d):
    # get_pad (networks/inpaintor.py:7-9) for even input sizes: ((ceil(n/s)-1)*s + d*(k-1) + 1 - n) / 2
    return {(5, 1, 1): 2, (4, 2, 1): 1}.get((k, s, d), d * (k - 1) // 2)


def gated(x, sd, p, k, s, d, act):
    pad = _pad(k, s, d)
    a = F.conv2d(x, sd[p + '.conv2d.weight'], sd[p + '.conv2d.bias'], stride=s, padding=pad, dilation=d)
    m = F.conv2d(x, sd[p + '.mask_conv2d.weight'], sd[p + '.mask_conv2d.bias'], stride=s, padding=pad, dilation=d)
    y = (F.leaky_relu(a, 0.2) if act else a) * torch.sigmoid(m)
    return F.batch_norm(y, sd[p + '.batch_norm2d.running_mean'], sd[p + '.batch_norm2d.running_var'],
                        sd[p + '.batch_norm2d.weight'], sd[p + '.batch_norm2d.bias'], training=False, eps=1e-5)


def seq(x, sd, name, spec):
    for i, (k, s, d, deconv, act) in enumerate(spec):
        if deconv:
            x = gated(F.interpolate(x, scale_factor=2), sd, '%s.%d.conv2d' % (name, i), k, s, d, act)
        else:
            x = gated(x, sd, '%s.%d' % (name, i), k, s, d, act)
    return x


def self_attention(x, sd, p):
    b, C, w, h = x.shape
    q = F.conv2d(x, sd[p + '.query_conv.weight'], sd[p + '.query_conv.bias']).view(b, -1, w * h).permute(0, 2, 1)
    k = F.conv2d(x, sd[p + '.key_conv.weight'], sd[p + '.key_conv.bias']).view(b, -1, w * h)
    att = torch.softmax(torch.bmm(q, k), dim=-1)
    v = F.conv2d(x, sd[p + '.value_conv.weight'], sd[p + '.value_conv.bias']).view(b, -1, w * h)
    out = torch.bmm(v, att.permute(0, 2, 1)).view(b, C, w, h)
    return sd[p + '.gamma'] * out + x


def forward(imgs, masks, sd):
    masked = imgs * (1 - masks) + masks
    coarse = torch.clamp(seq(torch.cat([masked, masks], 1), sd, 'coarse_net', COARSE), -1., 1.)
    masked = imgs * (1 - masks) + coarse * masks
    x = seq(torch.cat([masked, masks], 1), sd, 'refine_conv_net', REFINE)
    x = self_attention(x, sd, 'refine_attn')
    x = torch.clamp(seq(x, sd, 'refine_upsample_net', UPSAMPLE), -1., 1.)
    return coarse, x, x * masks + imgs * (1 - masks)
