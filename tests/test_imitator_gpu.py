"""End-to-end GPU parity of the Imitator mirror (models/imitator.py:82-189) against the oracle:
personalize -> chunked inference over several target frames, host in / host out."""
import numpy as np
import pytest
import torch

from impersonator_b200 import synthetic as S
from impersonator_b200.generator import ImpersonatorGenerator
from impersonator_b200.imitator import Imitator, SyntheticBodyModel, morph
from impersonator_b200.nmr import SMPLRenderer
from oracle import generator_ref as G, nmr_ref

pytestmark = pytest.mark.gpu


class Opt(object):
    image_size, batch_size, bg_model, repeat_num, cond_nc = 256, 2, "ORIGINAL", 6, 3
    bg_ks, ft_ks, front_warp, only_vis = 13, 3, False, False


def ref_morph(mask, ks, mode='erode'):            # utils/util.py:73-89 verbatim (F.conv2d with a box kernel)
    import torch.nn.functional as F
    pad = ks // 2
    kernel = torch.ones(1, 1, ks, ks)
    if mode == 'erode':
        out = F.conv2d(F.pad(mask, [pad] * 4, value=1.0), kernel)
        return (out == ks * ks).float()
    out = F.conv2d(F.pad(mask, [pad] * 4, value=0.0), kernel)
    return (out >= 1).float()


def test_imitator_inference_matches_oracle(cuda):
    torch.set_grad_enabled(False)
    size = 256
    v, f = S.uv_sphere()
    tabs = S.synthetic_tables()
    net = ImpersonatorGenerator(bg_dim=4, src_dim=6, tsf_dim=6, repeat_num=6)
    sd = S.fill_state_dict(net.state_dict(), seed=0)
    net.load_state_dict(sd)
    render = SMPLRenderer(image_size=size, faces=f.numpy(), map_fn=tabs["map_fn"])
    body = SyntheticBodyModel(v)
    im = Imitator(Opt(), generator=net, hmr=body, render=render, device=cuda)
    src_img = S.synthetic_source(size)
    src_theta = np.zeros(85, np.float32)
    src_theta[0], src_theta[3], src_theta[4] = 0.95, 0.3, 0.1
    im.personalize("", src_smpl=src_theta, src_img=src_img)
    g = torch.Generator().manual_seed(5)
    tgt = np.zeros((5, 85), np.float32)
    tgt[:, 0] = 0.8 + 0.3 * torch.rand(5, generator=g).numpy()
    tgt[:, 1:3] = (torch.rand(5, 2, generator=g).numpy() * 2 - 1) * 0.1
    tgt[:, 3] = (torch.rand(5, generator=g).numpy() * 2 - 1) * 3.0
    tgt[:, 4] = (torch.rand(5, generator=g).numpy() * 2 - 1) * 0.3
    outs = im.inference_by_smpls(list(tgt), cam_strategy="smooth")          # chunks of 2, 2, 1
    assert len(outs) == 5 and outs[0].shape == (size, size, 3) and outs[0].dtype == np.float32

    # ---- oracle: the reference's personalize + per-frame loop (models/imitator.py:82-189), on CPU
    sth = torch.from_numpy(src_theta)[None]
    sinfo = body.get_details(sth)
    f2v, sfim, _ = nmr_ref.render_fim_wim(sinfo["cam"], sinfo["verts"], f, size)
    cond = nmr_ref.encode_fim(sfim, tabs["map_fn"])
    p2v = nmr_ref.src_p2verts(f2v)
    bg_mask = ref_morph(cond[:, -1:], 13, 'erode')
    bg = G.resnet_generator(torch.cat([src_img * bg_mask, bg_mask], dim=1), sd, 'bg_model')
    ft_mask = 1 - ref_morph(cond[:, -1:], 3, 'erode')
    feats = G.encode_src(torch.cat([src_img * ft_mask, cond], dim=1), sd)
    assert torch.equal(morph(cond[:, -1:], 13), bg_mask)
    first_cam = torch.from_numpy(tgt[0:1, 0:3])
    worst = 0.0
    for t in range(5):
        th = torch.from_numpy(tgt[t:t + 1])
        cam = sinfo["cam"].clone()
        cam[:, 1:] += th[:, 1:3] - first_cam[:, 1:]                          # swap_smpl 'smooth' (:224-227)
        tsf = body.get_details(torch.cat([cam, th[:, 3:75], sinfo["shape"]], dim=1))
        c = nmr_ref.correspond(tsf["cam"], tsf["verts"], f, tabs["map_fn"], p2v, src_img, size)
        pred, _, _ = G.imitator_forward(bg, feats, c["tsf_inputs"], c["T"], sd)
        d = np.abs(outs[t] - pred[0].permute(1, 2, 0).numpy()).max()
        worst = max(worst, d)
    print("Imitator.inference_by_smpls vs oracle loop: max-abs %.3e over 5 frames" % worst)
    assert worst < 1e-3
    assert im.tsf_info["T"].shape[0] == 1                                    # tsf_info describes the last frame
