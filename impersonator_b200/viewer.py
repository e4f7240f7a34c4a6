"""Host-side mirror of models/viewer.py's ``Viewer`` (novel-view synthesis, inference only).

The reference's Viewer is its Imitator plus a rigid transform of the source body (models/viewer.py:237-279): rotate /
translate the source vertices, rasterize, compute the transformation flow against the source's visible faces, warp the
source image and run ``generator.inference`` on the source's cached features.  Everything below ``view`` is the hot path
this library already provides (one fused correspondence pass + the tcgen05 conv engine), so the class only adds

  rotate_trans(rt, t, X)                           models/viewer.py:237-244
  view(rt, t, visualizer=None, name='1')           :246-279   -> preds [N,3,H,W] on the device
  forward(tsf_inputs, feats, T, bg)                :281-288   -> (preds, tsf_mask)

``run_view.py:62-73`` calls ``view`` sixteen times, one angle each.  ``view`` also accepts ``rt`` / ``t`` of shape [N,3]
(extension): the N views then run as ONE batch, which is how the B200 path is meant to be driven (``view_many``).
personalize / inference / inference_by_smpls are the Imitator's (the reference classes share that code verbatim, except
for the ``--bg_model ORIGINAL`` background paste of models/viewer.py:129).  ``post_personalize`` is out of scope (backward).
"""
import numpy as np
import torch

from ._lib import LwbError
from .imitator import Imitator, _on_device


def euler2matrix(rt):
    """utils/cv_utils.py:333-353: R = Rz @ Ry @ Rx from Euler angles (radians), float32 like the reference."""
    rt = np.asarray(rt, dtype=np.float64)
    cx, sx = np.cos(rt[0]), np.sin(rt[0])
    cy, sy = np.cos(rt[1]), np.sin(rt[1])
    cz, sz = np.cos(rt[2]), np.sin(rt[2])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], dtype=np.float32)
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=np.float32)
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]], dtype=np.float32)
    return np.dot(Rz, np.dot(Ry, Rx))


class Viewer(Imitator):

    def __init__(self, opt, **kw):
        super(Viewer, self).__init__(opt, **kw)
        self._name = 'Viewer'

    def _original_bg(self, bg_inputs, img_bg):
        return bg_inputs[:, 0:3] + img_bg * bg_inputs[:, -1:]             # models/viewer.py:129

    def rotate_trans(self, rt, t, X):
        """X [1|N,V,3] -> X @ R + t (models/viewer.py:237-244); rt / t may be [3] or [N,3]."""
        rt = np.asarray(rt, dtype=np.float32).reshape(-1, 3)
        t = np.asarray(t, dtype=np.float32).reshape(-1, 3)
        n = max(rt.shape[0], t.shape[0])
        if rt.shape[0] not in (1, n) or t.shape[0] not in (1, n) or X.shape[0] not in (1, n):
            raise LwbError("rotate_trans: rt %r, t %r and X %r do not broadcast" % (rt.shape, t.shape, tuple(X.shape)))
        R = torch.from_numpy(np.stack([euler2matrix(r) for r in rt])).to(X.device)
        tt = torch.from_numpy(t).to(X.device)[:, None, :]
        return torch.matmul(X.expand(n, -1, -1) if X.shape[0] != n else X, R) + tt

    @_on_device
    @torch.no_grad()
    def view(self, rt, t, visualizer=None, name='1'):
        src_info = self.src_info
        tsf_mesh = self.rotate_trans(rt, t, src_info['verts']).contiguous()
        n = tsf_mesh.shape[0]
        cam = src_info['cam'].expand(n, -1).contiguous()
        out = self.render.correspond(cam, tsf_mesh, src_info['p2verts'], src_info['img'], align_corners=self._ac)
        bg = src_info['bg'] if getattr(self._opt, 'bg_replace', False) else torch.zeros_like(src_info['bg'])
        preds, tsf_mask = self.forward(out['tsf_inputs'], src_info['feats'], out['T'], bg)
        if getattr(self._opt, 'front_warp', False):
            preds = self.warp_front(preds, out['tsf_img'], out['fim'], tsf_mask)
        self.tsf_info = dict(fim=out['fim'], wim=out['wim'], cond=out['cond'], tsf_img=out['tsf_img'], T=out['T'],
                             cam=cam, verts=tsf_mesh)
        if visualizer is not None:
            visualizer.vis_named_img('src_img', src_info['img'])
            visualizer.vis_named_img('pred_' + name, preds)
            visualizer.vis_named_img('cond_' + name, out['cond'])
        return preds

    def view_many(self, rts, ts):
        """N views as one batch: rts [N,3] (radians), ts [N,3] or [3] -> preds [N,3,H,W]."""
        return self.view(np.asarray(rts, dtype=np.float32).reshape(-1, 3), ts)

    @_on_device
    @torch.no_grad()
    def forward(self, tsf_inputs, feats, T, bg):
        src_encoder_outs, src_resnet_outs = feats
        tsf_color, tsf_mask, pred_imgs = self.generator.inference(src_encoder_outs, src_resnet_outs, tsf_inputs, T, bg=bg)
        return pred_imgs, tsf_mask

    def warp_front(self, preds, tsf_img, fim, mask):
        front_mask = self.render.encode_front_fim(fim, transpose=True, front_fn=True)
        return (1 - front_mask) * preds + tsf_img * front_mask * (1 - mask)
