"""Host-side mirror of utils/nmr.py's ``SMPLRenderer`` for the inference hot path.

Same constructor arguments and method signatures (utils/nmr.py:104-107,263,328,343,506,617);
the per-frame work runs in the fused sm_100a correspondence kernels (lwb_correspond) instead of
>100 small ATen launches + the O(pixels x faces) rasterizer:

  render_fim_wim(cam, vertices, faces=None) -> (f2verts, fim, wim)          utils/nmr.py:263-278
  encode_fim(cam, vertices, fim=None, transpose=True, map_fn=None)          utils/nmr.py:328-341
  encode_front_fim(fim, transpose=True, front_fn=True)                      utils/nmr.py:343-352
  cal_bc_transform(src_f2pts, dst_fims, dst_wims)                           utils/nmr.py:617-659
  get_vis_f2pts(f2pts, fims)                                                utils/nmr.py:506-546
  correspond(cam, vertices, src_p2verts, src_img)  [new]  everything transfer_params_by_smpl needs
                                                           (models/imitator.py:251-260) in one pass

Construction follows utils/nmr.py:104-178: ``smpl_faces.npy`` + the lookup tables built from ``mapper.txt`` /
``front_facial.json`` / ``head.json`` by impersonator_b200.mesh (utils/mesh.py:368-421).  Those assets are external
downloads in the reference (README.md:48-68); the tables can also be passed in directly (``faces=``, ``map_fn=``,
...), which is how the synthetic benchmarks run.
Textured / lit rendering (``render``, ``extract_tex``, ...) is visualisation only: out of scope.
"""
import os

import numpy as np
import torch
import torch.nn as nn

from . import kernels as K
from . import neural_renderer as nr
from ._lib import LwbError


def orthographic_proj_withz_idrot(X, cam, offset_z=0.):
    """utils/nmr.py:10-28: weak perspective, xy = s * (XY + t), z kept."""
    scale = cam[:, 0].contiguous().view(-1, 1, 1)
    trans = cam[:, 1:3].contiguous().view(cam.size(0), 1, -1)
    proj_xy = scale * (X[:, :, :2] + trans)
    proj_z = X[:, :, 2, None] + offset_z
    return torch.cat((proj_xy, proj_z), 2)


class SMPLRenderer(nn.Module):
    def __init__(self, face_path='assets/pretrains/smpl_faces.npy',
                 uv_map_path='assets/pretrains/mapper.txt', map_name='uv_seg', tex_size=3, image_size=256,
                 anti_aliasing=True, fill_back=False, background_color=(0, 0, 0), viewing_angle=30, near=0.1, far=25.0,
                 has_front=False, faces=None, map_fn=None, back_map_fn=None, front_map_fn=None):
        super(SMPLRenderer, self).__init__()
        self.background_color = background_color
        self.anti_aliasing = anti_aliasing
        self.image_size = image_size
        self.fill_back = fill_back
        self.map_name = map_name
        self.tex_size = tex_size
        from_files = faces is None
        if from_files:
            # the reference's own construction path (utils/nmr.py:137-161): tables from the asset files
            if not os.path.exists(face_path) or not os.path.exists(uv_map_path):
                raise LwbError("%s / %s not found: the SMPL assets are an external download (README.md:48-68); "
                               "or pass faces= / map_fn= tables explicitly" % (face_path, uv_map_path))
            faces = np.load(face_path)
        faces = torch.as_tensor(np.asarray(faces).astype(np.int32)).int()
        self.base_nf = faces.shape[0]
        if self.fill_back:
            faces = torch.cat((faces, faces.flip(1)), dim=0)
        self.nf = faces.shape[0]
        self.register_buffer('faces', faces.contiguous())
        if from_files:
            from . import mesh
            if map_fn is None:
                map_fn = mesh.create_mapping(map_name, uv_map_path, contain_bg=True, fill_back=fill_back)
            if back_map_fn is None:
                back_map_fn = mesh.create_mapping('back', uv_map_path, contain_bg=True, fill_back=fill_back)
            if has_front and front_map_fn is None:
                front_map_fn = mesh.create_mapping('front', uv_map_path, contain_bg=True, fill_back=fill_back)
        if map_fn is None:
            raise LwbError("map_fn table required when faces= is given explicitly")
        if np.asarray(map_fn).shape[0] != self.nf + 1:
            raise LwbError("map_fn has %d rows, the mesh %d faces (+1 background row)" % (np.asarray(map_fn).shape[0], self.nf))
        self.register_buffer('map_fn', torch.as_tensor(np.asarray(map_fn)).float().contiguous())
        if back_map_fn is not None:
            self.register_buffer('back_map_fn', torch.as_tensor(np.asarray(back_map_fn)).float().contiguous())
        else:
            self.back_map_fn = None
        if has_front:
            if front_map_fn is None:
                raise LwbError("has_front=True needs front_map_fn")
            self.register_buffer('front_map_fn', torch.as_tensor(np.asarray(front_map_fn)).float().contiguous())
        else:
            self.front_map_fn = None
        self.rasterizer_eps = 1e-3
        self.near = near
        self.far = far
        self.proj_func = orthographic_proj_withz_idrot
        self.viewing_angle = viewing_angle
        self.eye = [0, 0, -(1. / np.tan(np.radians(self.viewing_angle)) + 1)]

    # ---- reference API --------------------------------------------------------------------
    @torch.no_grad()
    def render_fim_wim(self, cam, vertices, faces=None):
        """-> f2verts f32[B,F,3,3], fim i32[B,H,W], wim f32[B,H,W,3] (utils/nmr.py:263-278)."""
        if faces is not None:
            # explicit per-sample faces: generic path through the nr mirror
            proj_verts = self.proj_func(vertices, cam)
            proj_verts[:, :, 1] *= -1
            verts = nr.look_at(proj_verts, self.eye)
            f2v = nr.vertices_to_faces(verts, faces)
            fim, wim = nr.rasterize_face_index_map_and_weight_map(f2v, self.image_size, False)
            return f2v, fim, wim
        out = self._correspond(cam, vertices, None, None, want_f2verts=True)
        return out['f2verts'], out['fim'], out['wim']

    @torch.no_grad()
    def render_silhouettes(self, cam, vertices, faces=None):
        f2v, _, _ = self.render_fim_wim(cam, vertices, faces)
        return nr.rasterize_silhouettes(f2v, self.image_size, self.anti_aliasing)

    @torch.no_grad()
    def encode_fim(self, cam, vertices, fim=None, transpose=True, map_fn=None):
        if fim is None:
            raise NotImplementedError                      # utils/nmr.py:311 infer_face_index_map raises too
        table = map_fn if map_fn is not None else self.map_fn
        fim_enc = table[fim.long()]                        # -1 -> last (background) row, utils/nmr.py:336
        if transpose:
            fim_enc = fim_enc.permute(0, 3, 1, 2)
        return fim_enc, fim

    @torch.no_grad()
    def encode_front_fim(self, fim, transpose=True, front_fn=True):
        table = self.front_map_fn if front_fn else self.back_map_fn
        fim_enc = table[fim.long()]
        if transpose:
            fim_enc = fim_enc.permute(0, 3, 1, 2)
        return fim_enc

    @torch.no_grad()
    def cal_bc_transform(self, src_f2pts, dst_fims, dst_wims):
        """T[b,p] = sum_k wim[b,p,k] * src_f2pts[b, fim[b,p], k, :], -2 where uncovered (utils/nmr.py:617-659).
        Standalone (cold) form; the per-frame path gets T from correspond()."""
        bs = src_f2pts.shape[0]
        idx = dst_fims.long().reshape(bs, -1)
        mask = idx >= 0
        pts = torch.gather(src_f2pts.reshape(bs, -1, 6), 1, idx.clamp(min=0)[:, :, None].expand(-1, -1, 6)).view(bs, -1, 3, 2)
        T = (pts * dst_wims.reshape(bs, -1, 3)[:, :, :, None]).sum(dim=2)
        T = torch.where(mask[:, :, None], T, torch.full_like(T, -2.0))
        return T.view(bs, self.image_size, self.image_size, 2)

    @staticmethod
    def create_meshgrid(image_size):
        """utils/nmr.py:491-504: the identity sampling grid [H,W,2] in [-1,1], (x, y) order."""
        factor = (torch.arange(0, image_size, dtype=torch.float32) / (image_size - 1) - 0.5) * 2
        xv, yv = torch.meshgrid([factor, factor], indexing='ij')
        return torch.stack([yv, xv], dim=-1)

    @staticmethod
    def get_vis_f2pts(f2pts, fims):
        """utils/nmr.py:506-546: keep visible faces' points, -2 elsewhere."""
        def get_vis(orig, fim):
            vis = torch.zeros_like(orig) - 2.0
            ids = fim.unique()[1:].long()                 # utils/nmr.py:528: the first unique value is taken to be -1
            vis[ids] = orig[ids]
            return vis
        if f2pts.dim() == 4:
            return torch.stack([get_vis(f2pts[i], fims[i]) for i in range(f2pts.shape[0])], dim=0)
        return get_vis(f2pts, fims)

    # ---- fused per-frame path -------------------------------------------------------------
    def _correspond(self, cam, vertices, src_p2verts, src_img, want_f2verts=False, align_corners=None):
        if not vertices.is_cuda:
            raise LwbError("SMPLRenderer runs on CUDA tensors only (no CPU fallback)")
        cam = cam.float().contiguous()
        vertices = vertices.float().contiguous()
        if src_p2verts is None:
            # no correspondence target: any valid table (T is then meaningless and ignored)
            if getattr(self, '_dummy_p2v', None) is None or self._dummy_p2v.device != vertices.device:
                self._dummy_p2v = torch.zeros((1, self.nf, 3, 2), dtype=torch.float32, device=vertices.device)
            src_p2verts = self._dummy_p2v
        return K.correspond(cam, vertices, self.faces, self.image_size, self.map_fn, src_p2verts.float().contiguous(),
                            src_img.float().contiguous() if src_img is not None else None,
                            align_corners=align_corners, want_f2verts=want_f2verts)

    @torch.no_grad()
    def correspond(self, cam, vertices, src_p2verts, src_img=None, align_corners=None, want_f2verts=False):
        """One pass = render_fim_wim + encode_fim + cal_bc_transform + F.grid_sample(src_img, T) + cat
        (models/imitator.py:251-260).  src_p2verts [1|B,F,3,2], src_img [1|B,3,H,W].
        -> dict(fim, wim, cond, T, tsf_img, tsf_inputs[, f2verts])."""
        return self._correspond(cam, vertices, src_p2verts, src_img, want_f2verts, align_corners)

    def render(self, *a, **k):
        raise LwbError("textured rendering is visualisation only and outside the hot path (SURVEY.md section 8)")

    extract_tex = render
    extract_tex_from_image = render
