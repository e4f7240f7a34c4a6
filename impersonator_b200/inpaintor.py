"""Host-side mirror of networks/inpaintor.py (DeepFill-v2 style ``InpaintSANet``), the default
background network of ``Imitator.personalize`` (models/imitator.py:48-52,124-125; once per source).

Same module tree and ``state_dict`` keys as the reference (322 tensors incl. BatchNorm running
stats): ``GatedConv2dWithActivation`` (conv2d, mask_conv2d, batch_norm2d), ``GatedDeConv2dWithActivation``
(nearest 2x + gated conv), ``SelfAttention`` (query/key/value 1x1 convs, gamma), ``InpaintSANet``
(coarse_net 17 / refine_conv_net 11 / refine_attn / refine_upsample_net 7).

Execution: every gated layer = ONE direct-conv launch computing both convs (weights stacked on the
output channels) + one fused gate/BatchNorm kernel (lwb_conv2d_direct_nchw, lwb_gated_bn_nchw).
This path runs once per source image, off the per-frame loop; the 4096x4096 self-attention uses the
library batched GEMM (torch.bmm -> cuBLAS) + softmax, as a plain library call.
"""
import numpy as np
import torch
import torch.nn as nn

from . import kernels as K
from ._lib import LwbError


def get_pad(in_, ksize, stride, atrous=1):
    """networks/inpaintor.py:7-9."""
    out_ = np.ceil(float(in_) / stride)
    return int(((out_ - 1) * stride + atrous * (ksize - 1) + 1 - in_) / 2)


class GatedConv2dWithActivation(nn.Module):
    """phi(f(I)) * sigmoid(g(I)), then BatchNorm (networks/inpaintor.py:12-47)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 batch_norm=True, activation=torch.nn.LeakyReLU(0.2, inplace=True)):
        super(GatedConv2dWithActivation, self).__init__()
        if groups != 1:
            raise LwbError("grouped gated convs are not used by InpaintSANet")
        self.batch_norm = batch_norm
        self.activation = activation
        self.conv2d = torch.nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        self.mask_conv2d = torch.nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        self.batch_norm2d = torch.nn.BatchNorm2d(out_channels)
        self.sigmoid = torch.nn.Sigmoid()
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
        self._packed = None

    def _pack(self):
        dev = self.conv2d.weight.device
        if self._packed is None or self._packed[0].device != dev:
            w = torch.cat([self.conv2d.weight, self.mask_conv2d.weight], dim=0).detach().float().contiguous()
            b = None
            if self.conv2d.bias is not None:
                b = torch.cat([self.conv2d.bias, self.mask_conv2d.bias]).detach().float().contiguous()
            scale = shift = None
            if self.batch_norm:
                bn = self.batch_norm2d
                scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach().float().contiguous()
                shift = (bn.bias - bn.running_mean * scale).detach().float().contiguous()
            self._packed = (w, b, scale, shift)
        return self._packed

    @torch.no_grad()
    def forward(self, input):
        if self.training:
            raise LwbError("the B200 path is inference-only (eval-mode BatchNorm)")
        w, b, scale, shift = self._pack()
        ab = K.conv2d_direct_nchw(input.float().contiguous(), w, b, stride=self.conv2d.stride[0],
                                  pad=self.conv2d.padding[0], dil=self.conv2d.dilation[0])
        act = 0 if self.activation is None else 2
        return K.gated_bn_nchw(ab, act, scale, shift)


class GatedDeConv2dWithActivation(nn.Module):
    """nearest 2x resize + gated conv (networks/inpaintor.py:50-68)."""

    def __init__(self, scale_factor, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, batch_norm=True, activation=torch.nn.LeakyReLU(0.2, inplace=True)):
        super(GatedDeConv2dWithActivation, self).__init__()
        self.conv2d = GatedConv2dWithActivation(in_channels, out_channels, kernel_size, stride, padding, dilation, groups,
                                                bias, batch_norm, activation)
        self.scale_factor = scale_factor

    @torch.no_grad()
    def forward(self, input):
        n, c, h, w = input.shape
        x = input[:, :, :, None, :, None].expand(n, c, h, 2, w, 2).reshape(n, c, 2 * h, 2 * w)   # nearest x2 (:67)
        return self.conv2d(x)


class SelfAttention(nn.Module):
    """networks/inpaintor.py:71-107."""

    def __init__(self, in_dim, activation, with_attn=False):
        super(SelfAttention, self).__init__()
        self.chanel_in = in_dim
        self.activation = activation
        self.with_attn = with_attn
        self.query_conv = nn.Conv2d(in_channels=in_dim, out_channels=in_dim // 8, kernel_size=1)
        self.key_conv = nn.Conv2d(in_channels=in_dim, out_channels=in_dim // 8, kernel_size=1)
        self.value_conv = nn.Conv2d(in_channels=in_dim, out_channels=in_dim, kernel_size=1)
        self.gamma = nn.Parameter(torch.zeros(1))
        self.softmax = nn.Softmax(dim=-1)

    @torch.no_grad()
    def forward(self, x):
        b, C, width, height = x.size()
        x = x.float().contiguous()
        conv = lambda m: K.conv2d_direct_nchw(x, m.weight.detach().float().contiguous(), m.bias.detach().float().contiguous())
        q = conv(self.query_conv).view(b, -1, width * height).permute(0, 2, 1)
        k = conv(self.key_conv).view(b, -1, width * height)
        attention = torch.softmax(torch.bmm(q, k), dim=-1)                   # library GEMM + softmax (cold path)
        v = conv(self.value_conv).view(b, -1, width * height)
        out = torch.bmm(v, attention.permute(0, 2, 1)).view(b, C, width, height)
        out = self.gamma * out + x
        return (out, attention) if self.with_attn else out


class InpaintSANet(torch.nn.Module):
    """networks/inpaintor.py:110-202."""

    def __init__(self, c_dim=5):
        super(InpaintSANet, self).__init__()
        cnum = 32
        G, D = GatedConv2dWithActivation, GatedDeConv2dWithActivation
        self.coarse_net = nn.Sequential(
            G(c_dim, cnum, 5, 1, padding=get_pad(256, 5, 1)),
            G(cnum, 2 * cnum, 4, 2, padding=get_pad(256, 4, 2)),
            G(2 * cnum, 2 * cnum, 3, 1, padding=get_pad(128, 3, 1)),
            G(2 * cnum, 4 * cnum, 4, 2, padding=get_pad(128, 4, 2)),
            G(4 * cnum, 4 * cnum, 3, 1, padding=get_pad(64, 3, 1)),
            G(4 * cnum, 4 * cnum, 3, 1, padding=get_pad(64, 3, 1)),
            G(4 * cnum, 4 * cnum, 3, 1, dilation=2, padding=get_pad(64, 3, 1, 2)),
            G(4 * cnum, 4 * cnum, 3, 1, dilation=4, padding=get_pad(64, 3, 1, 4)),
            G(4 * cnum, 4 * cnum, 3, 1, dilation=8, padding=get_pad(64, 3, 1, 8)),
            G(4 * cnum, 4 * cnum, 3, 1, dilation=16, padding=get_pad(64, 3, 1, 16)),
            G(4 * cnum, 4 * cnum, 3, 1, padding=get_pad(64, 3, 1)),
            G(4 * cnum, 4 * cnum, 3, 1, padding=get_pad(64, 3, 1)),
            D(2, 4 * cnum, 2 * cnum, 3, 1, padding=get_pad(128, 3, 1)),
            G(2 * cnum, 2 * cnum, 3, 1, padding=get_pad(128, 3, 1)),
            D(2, 2 * cnum, cnum, 3, 1, padding=get_pad(256, 3, 1)),
            G(cnum, cnum // 2, 3, 1, padding=get_pad(256, 3, 1)),
            G(cnum // 2, 3, 3, 1, padding=get_pad(128, 3, 1), activation=None))
        self.refine_conv_net = nn.Sequential(
            G(c_dim, cnum, 5, 1, padding=get_pad(256, 5, 1)),
            G(cnum, cnum, 4, 2, padding=get_pad(256, 4, 2)),
            G(cnum, 2 * cnum, 3, 1, padding=get_pad(128, 3, 1)),
            G(2 * cnum, 2 * cnum, 4, 2, padding=get_pad(128, 4, 2)),
            G(2 * cnum, 4 * cnum, 3, 1, padding=get_pad(64, 3, 1)),
            G(4 * cnum, 4 * cnum, 3, 1, padding=get_pad(64, 3, 1)),
            G(4 * cnum, 4 * cnum, 3, 1, padding=get_pad(64, 3, 1)),
            G(4 * cnum, 4 * cnum, 3, 1, dilation=2, padding=get_pad(64, 3, 1, 2)),
            G(4 * cnum, 4 * cnum, 3, 1, dilation=4, padding=get_pad(64, 3, 1, 4)),
            G(4 * cnum, 4 * cnum, 3, 1, dilation=8, padding=get_pad(64, 3, 1, 8)),
            G(4 * cnum, 4 * cnum, 3, 1, dilation=16, padding=get_pad(64, 3, 1, 16)))
        self.refine_attn = SelfAttention(4 * cnum, 'relu', with_attn=False)
        self.refine_upsample_net = nn.Sequential(
            G(4 * cnum, 4 * cnum, 3, 1, padding=get_pad(64, 3, 1)),
            G(4 * cnum, 4 * cnum, 3, 1, padding=get_pad(64, 3, 1)),
            D(2, 4 * cnum, 2 * cnum, 3, 1, padding=get_pad(128, 3, 1)),
            G(2 * cnum, 2 * cnum, 3, 1, padding=get_pad(128, 3, 1)),
            D(2, 2 * cnum, cnum, 3, 1, padding=get_pad(256, 3, 1)),
            G(cnum, cnum // 2, 3, 1, padding=get_pad(256, 3, 1)),
            G(cnum // 2, 3, 3, 1, padding=get_pad(256, 3, 1), activation=None))

    @torch.no_grad()
    def forward(self, imgs, masks, only_out=False, only_x=False):
        masked_imgs = imgs * (1 - masks) + masks
        x = self.coarse_net(torch.cat([masked_imgs, masks], dim=1))
        coarse_x = torch.clamp(x, -1., 1.)
        masked_imgs = imgs * (1 - masks) + coarse_x * masks
        x = self.refine_conv_net(torch.cat([masked_imgs, masks], dim=1))
        x = self.refine_attn(x)
        x = self.refine_upsample_net(x)
        x = torch.clamp(x, -1., 1.)
        comp_imgs = x * masks + imgs * (1 - masks)
        if only_out:
            return comp_imgs
        if only_x:
            return x
        return coarse_x, x, comp_imgs
