"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the task classes next to the Imitator.

  personalize    models/viewer.py:83-143, models/swapper.py:99-165 (and models/imitator.py:82-145)
  Viewer.view    models/viewer.py:237-288  (rotate_trans :237-244, forward :281-288, warp_front :231-235)
  Swapper.swap   models/swapper.py:199-270 (calculate_trans :242-253, forward :261-270, warp :255-259)
  euler2matrix   utils/cv_utils.py:333-353

Composed from the pinned pieces (oracle/nmr_ref.py, oracle/generator_ref.py, oracle/raster.py); torch fp32 on CPU.
Pinned to the reference's code by tests/golden/tasks.npz: make_tasks_golden.py runs the reference's OWN
``models/viewer.py`` / ``models/swapper.py`` methods (unbound, CPU) on the same synthetic inputs, and
tests/test_tasks_cpu.py checks this file against those slices.  The product never imports it.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import generator_ref as G, nmr_ref


def euler2matrix(rt):
    rx, ry, rz = [float(a) for a in rt]
    Rx = np.array([[1, 0, 0], [0, np.cos(rx), -np.sin(rx)], [0, np.sin(rx), np.cos(rx)]], dtype=np.float32)
    Ry = np.array([[np.cos(ry), 0, np.sin(ry)], [0, 1, 0], [-np.sin(ry), 0, np.cos(ry)]], dtype=np.float32)
    Rz = np.array([[np.cos(rz), -np.sin(rz), 0], [np.sin(rz), np.cos(rz), 0], [0, 0, 1]], dtype=np.float32)
    return np.dot(Rz, np.dot(Ry, Rx))


def morph(mask, ks, mode='erode'):                       # utils/util.py:73-89
    pad = ks // 2
    kernel = torch.ones(1, 1, ks, ks)
    if mode == 'erode':
        return (F.conv2d(F.pad(mask, [pad] * 4, value=1.0), kernel) == ks * ks).float()
    return (F.conv2d(F.pad(mask, [pad] * 4, value=0.0), kernel) >= 1).float()


def personalize(img, cam, verts, faces, tabs, sd, size, task, part_fn=None, bg_ks=13, ft_ks=3, only_vis=False):
    """img [1,3,H,W] in [-1,1]; cam [1,3], verts [1,V,3] -> src_info.  ``task`` in {'imitator','viewer','swapper'} selects
    what is kept of the ORIGINAL background net's output (viewer.py:129 vs imitator.py:131 / swapper.py:148)."""
    f2v, fim, wim = nmr_ref.render_fim_wim(cam, verts, faces, size)
    cond = nmr_ref.encode_fim(fim, tabs["map_fn"])
    info = dict(cam=cam, verts=verts, fim=fim, wim=wim, cond=cond, f2verts=f2v, p2verts=nmr_ref.src_p2verts(f2v), img=img)
    if only_vis:                                          # models/imitator.py:109-110: hidden faces' points -> -2
        info['p2verts'] = nmr_ref.get_vis_f2pts(info['p2verts'], fim)
    if part_fn is not None:
        info['part'] = nmr_ref.encode_fim(fim, part_fn)
    bg_mask = morph(cond[:, -1:], bg_ks, 'erode')
    bg_inputs = torch.cat([img * bg_mask, bg_mask], dim=1)
    img_bg = G.resnet_generator(bg_inputs, sd, 'bg_model')
    info['bg'] = bg_inputs[:, 0:3] + img_bg * bg_inputs[:, -1:] if task == 'viewer' else img_bg
    ft_mask = 1 - morph(cond[:, -1:], ft_ks, 'erode')
    info['src_inputs'] = torch.cat([img * ft_mask, cond], dim=1)
    info['feats'] = G.encode_src(info['src_inputs'], sd)
    return info


def view(src_info, rt, t, faces, tabs, sd, size, bg_replace=False, front_warp=False):
    R = torch.from_numpy(euler2matrix(rt))[None]
    mesh = torch.bmm(src_info['verts'], R) + torch.as_tensor(t, dtype=torch.float32)[None, None, :]
    c = nmr_ref.correspond(src_info['cam'], mesh, faces, tabs["map_fn"], src_info['p2verts'], src_info['img'], size)
    bg = src_info['bg'] if bg_replace else torch.zeros_like(src_info['bg'])
    preds, _, mask = G.imitator_forward(bg, src_info['feats'], c['tsf_inputs'], c['T'], sd)
    if front_warp:
        front = nmr_ref.encode_fim(c['fim'], tabs["front_map_fn"])
        preds = (1 - front) * preds + c['tsf_img'] * front * (1 - mask)
    return preds, mesh


def calculate_trans(src_info, tgt_info, left_mask, left_faces, size):
    factor = (torch.arange(0, size, dtype=torch.float32) / (size - 1) - 0.5) * 2          # utils/nmr.py:499-503
    xv, yv = torch.meshgrid([factor, factor], indexing='ij')
    T11 = torch.stack([yv, xv], dim=-1)
    T11[~left_mask[0]] = -2
    f2p = tgt_info['p2verts'].clone()
    f2p[0, left_faces] = -2
    T21 = nmr_ref.cal_bc_transform(f2p, src_info['fim'], src_info['wim'], size).clamp(-2, 2)
    return T11[None], T21


def swap(src_info, tgt_info, part_faces, tabs, sd, size, target_part='body', front_warp=False):
    ids = {'body': list(range(1, 10)), 'all': list(range(10))}[target_part]
    left_ids = [i for i in range(10) if i not in ids]
    part_mask = (src_info['part'][:, ids].sum(dim=1) != 0)
    left_mask = src_info['part'][:, left_ids].sum(dim=1).bool() if left_ids else torch.zeros_like(part_mask)
    left_faces = sorted(set().union(*[set(part_faces[i]) for i in left_ids])) if left_ids else []
    T11, T21 = calculate_trans(src_info, tgt_info, left_mask, left_faces, size)
    tsf21 = nmr_ref.grid_sample(tgt_info['img'], T21)
    tsf11 = nmr_ref.grid_sample(src_info['img'], T11)
    tsf_img = tsf21 * part_mask[:, None].float() + tsf11 * left_mask[:, None].float()
    tsf_inputs = torch.cat([tsf_img, src_info['cond']], dim=1)
    e21, r21 = tgt_info['feats']
    e11, r11 = src_info['feats']
    color, mask = G.swap(tsf_inputs, e21, e11, r21, r11, T21, T11, sd)
    preds = mask * src_info['bg'] + (1 - mask) * color
    if front_warp:
        front = nmr_ref.encode_fim(src_info['fim'], tabs["front_map_fn"])
        preds = (1 - front) * preds + src_info['img'] * front * (1 - mask)
    return preds, T11, T21


def imitate(src_info, src_shape, thetas, body, faces, tabs, sd, size, cam_strategy='smooth', front_warp=False):
    """Imitator.inference_by_smpls (models/imitator.py:192-268, 326-342): per frame swap_smpl -> body model -> raster +
    correspondence -> generator.inference + composite (+ warp_front).  thetas [N,85] -> list of HxWx3 arrays, last T."""
    outs, first_cam, T = [], None, None
    for t in range(thetas.shape[0]):
        th = thetas[t:t + 1]
        if t == 0 and cam_strategy == 'smooth':
            first_cam = th[:, 0:3].clone()
        if cam_strategy == 'smooth':
            cam = src_info['cam'].clone()
            cam[:, 1:] += th[:, 1:3] - first_cam[:, 1:]
        elif cam_strategy == 'source':
            cam = src_info['cam']
        else:
            cam = th[:, 0:3]
        d = body.get_details(torch.cat([cam, th[:, 3:75], src_shape], dim=1))
        c = nmr_ref.correspond(d['cam'], d['verts'], faces, tabs["map_fn"], src_info['p2verts'], src_info['img'], size)
        preds, _, mask = G.imitator_forward(src_info['bg'], src_info['feats'], c['tsf_inputs'], c['T'], sd)
        if front_warp:
            front = nmr_ref.encode_fim(c['fim'], tabs["front_map_fn"])
            preds = (1 - front) * preds + c['tsf_img'] * front * (1 - mask)
        outs.append(preds[0].permute(1, 2, 0).numpy())
        T = c['T']
    return outs, T
