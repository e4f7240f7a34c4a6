"""TEST INFRASTRUCTURE ONLY -- functional torch-fp32 restatement of networks/generator.py.

Takes the reference's ``state_dict`` (same keys) and evaluates with torch.nn.functional on CPU:
  ResidualBlock            networks/generator.py:8-20
  ResNetGenerator          :23-65        (bg_model, ``model.{i}`` keys)
  ResUnetGenerator         :68-184       (encoders / resnets / decoders / skippers / img_reg / attetion_reg)
  ImpersonatorGenerator    :187-320      (forward, encode_src, infer_front, inference, swap, LWB = resize_trans+stn)
  Imitator.forward         models/imitator.py:326-336 (composite)
The arithmetic lives in the installed torch 2.11 (conv2d / conv_transpose2d / instance_norm /
grid_sampler_2d / upsample_bilinear2d) -- the same library the reference modules call, so this
restatement is validated by direct comparison with the imported reference modules
(tests/golden/make_generator_golden.py, run where /root/reference exists) and by the committed
golden slices tests/golden/generator*.npz.  Tolerance vs the CUDA path: 1e-3 max-abs (BASELINE.json).

``align_corners`` of the LWB's F.grid_sample (networks/generator.py:313, called without the flag):
True = the reference's pinned torch 1.2 semantics (default here and in the product); False = what the
installed torch 2.11 does for the same call (opt-in, LWB_ALIGN_CORNERS=0 on the product side).
"""
import torch
import torch.nn.functional as F


def _inorm(x, sd, p):
    return F.instance_norm(x, weight=sd[p + '.weight'], bias=sd[p + '.bias'], eps=1e-5)


def _cir(x, sd, p, stride=1, padding=1):           # Conv + IN + ReLU  (Sequential idx 0,1,2)
    return F.relu(_inorm(F.conv2d(x, sd[p + '.0.weight'], stride=stride, padding=padding), sd, p + '.1'))


def residual_block(x, sd, p):                       # generator.py:8-20
    h = F.conv2d(x, sd[p + '.main.0.weight'], padding=1)
    h = F.relu(_inorm(h, sd, p + '.main.1'))
    h = F.conv2d(h, sd[p + '.main.3.weight'], padding=1)
    h = _inorm(h, sd, p + '.main.4')
    return x + h


def resnet_generator(x, sd, p, repeat_num=6, n_down=3):     # generator.py:23-65
    i = 0
    x = F.relu(_inorm(F.conv2d(x, sd['%s.model.%d.weight' % (p, i)], padding=3), sd, '%s.model.%d' % (p, i + 1)))
    i += 3
    for _ in range(n_down):
        x = F.relu(_inorm(F.conv2d(x, sd['%s.model.%d.weight' % (p, i)], stride=2, padding=1), sd, '%s.model.%d' % (p, i + 1)))
        i += 3
    for _ in range(repeat_num):
        x = residual_block(x, sd, '%s.model.%d' % (p, i))
        i += 1
    for _ in range(n_down):
        x = F.conv_transpose2d(x, sd['%s.model.%d.weight' % (p, i)], stride=2, padding=1, output_padding=1)
        x = F.relu(_inorm(x, sd, '%s.model.%d' % (p, i + 1)))
        i += 3
    x = F.conv2d(x, sd['%s.model.%d.weight' % (p, i)], padding=3)
    return torch.tanh(x)


def unet_encoder(x, sd, p, i):                      # generator.py:77-95
    if i == 0:
        return _cir(x, sd, '%s.encoders.0' % p, stride=1, padding=3)
    return _cir(x, sd, '%s.encoders.%d' % (p, i), stride=2, padding=1)


def unet_decode(x, encoder_outs, sd, p, n_down=3):  # generator.py:173-181
    d = x
    for i in range(n_down):
        d = F.conv_transpose2d(d, sd['%s.decoders.%d.0.weight' % (p, i)], stride=2, padding=1, output_padding=1)
        d = F.relu(_inorm(d, sd, '%s.decoders.%d.1' % (p, i)))
        skip = encoder_outs[n_down - 1 - i]
        d = torch.cat([skip, d], dim=1)
        d = _cir(d, sd, '%s.skippers.%d' % (p, i), stride=1, padding=1)
    return d


def unet_regress(x, sd, p):                         # generator.py:183-184
    img = torch.tanh(F.conv2d(x, sd[p + '.img_reg.0.weight'], padding=3))
    mask = torch.sigmoid(F.conv2d(x, sd[p + '.attetion_reg.0.weight'], padding=3))
    return img, mask


def unet_inference(x, sd, p, repeat_num=6, n_down=3):       # generator.py:136-147 (encode_src)
    outs = [unet_encoder(x, sd, p, 0)]
    for i in range(1, n_down + 1):
        outs.append(unet_encoder(outs[-1], sd, p, i))
    res, h = [], outs[-1]
    for i in range(repeat_num):
        h = residual_block(h, sd, '%s.resnets.%d' % (p, i))
        res.append(h)
    return outs, res


def resize_trans(x, T):                             # generator.py:303-311
    h, w = x.shape[2:]
    Ts = F.interpolate(T.permute(0, 3, 1, 2), size=(h, w), mode='bilinear', align_corners=True)
    return Ts.permute(0, 2, 3, 1)


def stn(x, T, align_corners=True):                 # generator.py:312-315
    if x.shape[0] != T.shape[0]:
        x = x.expand(T.shape[0], -1, -1, -1)
    return F.grid_sample(x, T, mode='bilinear', padding_mode='zeros', align_corners=align_corners)


def transform(x, T, align_corners=True):           # generator.py:317-320
    return stn(x, resize_trans(x, T), align_corners)


def encode_src(src_inputs, sd, repeat_num=6):       # generator.py:213-214
    return unet_inference(src_inputs, sd, 'src_model', repeat_num)


def inference(src_encoder_outs, src_resnet_outs, tsf_inputs, T, sd, repeat_num=6, n_down=3,
              align_corners=True):                  # generator.py:277-301
    tsf_x = unet_encoder(tsf_inputs, sd, 'tsf_model', 0)
    tsf_encoder_outs = [tsf_x]
    for i in range(1, n_down + 1):
        src_x = src_encoder_outs[i]
        warp = transform(src_x, T, align_corners)
        tsf_x = unet_encoder(tsf_x, sd, 'tsf_model', i) + warp
        tsf_encoder_outs.append(tsf_x)
    T_scale = resize_trans(src_x, T)
    for i in range(repeat_num):
        warp = stn(src_resnet_outs[i], T_scale, align_corners)
        tsf_x = residual_block(tsf_x, sd, 'tsf_model.resnets.%d' % i) + warp
    return unet_regress(unet_decode(tsf_x, tsf_encoder_outs, sd, 'tsf_model', n_down), sd, 'tsf_model')


def infer_front(src_inputs, tsf_inputs, T, sd, repeat_num=6, n_down=3, align_corners=True):   # :216-243
    src_x = unet_encoder(src_inputs, sd, 'src_model', 0)
    tsf_x = unet_encoder(tsf_inputs, sd, 'tsf_model', 0)
    src_outs, tsf_outs = [src_x], [tsf_x]
    for i in range(1, n_down + 1):
        src_x = unet_encoder(src_x, sd, 'src_model', i)
        warp = transform(src_x, T, align_corners)
        tsf_x = unet_encoder(tsf_x, sd, 'tsf_model', i) + warp
        src_outs.append(src_x)
        tsf_outs.append(tsf_x)
    T_scale = resize_trans(src_x, T)
    for i in range(repeat_num):
        src_x = residual_block(src_x, sd, 'src_model.resnets.%d' % i)
        warp = stn(src_x, T_scale, align_corners)
        tsf_x = residual_block(tsf_x, sd, 'tsf_model.resnets.%d' % i) + warp
    src_img, src_mask = unet_regress(unet_decode(src_x, src_outs, sd, 'src_model', n_down), sd, 'src_model')
    tsf_img, tsf_mask = unet_regress(unet_decode(tsf_x, tsf_outs, sd, 'tsf_model', n_down), sd, 'tsf_model')
    return src_img, src_mask, tsf_img, tsf_mask


def forward(bg_inputs, src_inputs, tsf_inputs, T, sd, repeat_num=6, align_corners=True):      # :204-211
    img_bg = resnet_generator(bg_inputs, sd, 'bg_model', repeat_num, 3)
    return (img_bg,) + infer_front(src_inputs, tsf_inputs, T, sd, repeat_num, 3, align_corners)


def swap(tsf_inputs, enc12, enc21, res12, res21, T12, T21, sd, repeat_num=6, n_down=3,
         align_corners=True):                       # generator.py:245-275
    tsf_x = unet_encoder(tsf_inputs, sd, 'tsf_model', 0)
    outs = [tsf_x]
    for i in range(1, n_down + 1):
        tsf_x = unet_encoder(tsf_x, sd, 'tsf_model', i) + transform(enc12[i], T12, align_corners) \
            + transform(enc21[i], T21, align_corners)
        outs.append(tsf_x)
    Ts12, Ts21 = resize_trans(enc12[-1], T12), resize_trans(enc21[-1], T21)
    for i in range(repeat_num):
        tsf_x = residual_block(tsf_x, sd, 'tsf_model.resnets.%d' % i) + stn(res12[i], Ts12, align_corners) \
            + stn(res21[i], Ts21, align_corners)
    return unet_regress(unet_decode(tsf_x, outs, sd, 'tsf_model', n_down), sd, 'tsf_model')


def imitator_forward(bg_img, src_feats, tsf_inputs, T, sd, repeat_num=6, align_corners=True):
    """Imitator.forward (models/imitator.py:326-336, front_warp off)."""
    enc, res = src_feats
    color, mask = inference(enc, res, tsf_inputs, T, sd, repeat_num, 3, align_corners)
    return mask * bg_img + (1 - mask) * color, color, mask
