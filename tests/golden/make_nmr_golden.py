"""Generates tests/golden/nmr.npz -- run ONLY in the build container (needs /root/reference).

Pins the geometry glue of the hot path to the reference's OWN code: ``utils/nmr.py`` is imported
unmodified from /root/reference and its ``SMPLRenderer`` methods are called as unbound functions
on a namespace that carries the few attributes they read (the constructor itself needs the
external ``assets/pretrains/*`` files and ``.cuda()``):

  SMPLRenderer.render_fim_wim     utils/nmr.py:263-278   (proj :269, y flip :271, look_at :273, gather :276)
  SMPLRenderer.encode_fim         utils/nmr.py:328-341
  SMPLRenderer.encode_front_fim   utils/nmr.py:343-352
  SMPLRenderer.get_vis_f2pts      utils/nmr.py:506-546
  SMPLRenderer.cal_bc_transform   utils/nmr.py:617-659
  F.grid_sample(src_img, T) + cat (models/imitator.py:259-260; torch-1.2 semantics = align_corners=True)

``import neural_renderer as nr`` (utils/nmr.py:6) is satisfied by a stub module whose ``look_at`` /
``vertices_to_faces`` are the reference's own files (pure torch) and whose
``rasterize_face_index_map_and_weight_map`` is the CPU rasterizer oracle (oracle/raster.py: the C
restatement pinned to the reference's teapot goldens, bit-identical to the reference CUDA kernels
on the GPU box) -- the reference's rasterizer is CUDA-only and cannot run in this container.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from impersonator_b200 import synthetic as S          # noqa: E402
from oracle import nmr_ref, raster                    # noqa: E402


def _load_nr(name):
    spec = importlib.util.spec_from_file_location(
        "nr_" + name, os.path.join(REF, "thirdparty/neural_renderer/neural_renderer", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _rasterize(faces, image_size, anti_aliasing):
    assert anti_aliasing is False
    fim, wim, _ = raster.rasterize_fim_wim(faces.numpy(), image_size)
    return torch.from_numpy(fim), torch.from_numpy(wim)


def import_reference_nmr():
    nr = types.ModuleType("neural_renderer")
    nr.look_at = _load_nr("look_at").look_at
    nr.vertices_to_faces = _load_nr("vertices_to_faces").vertices_to_faces
    nr.rasterize_face_index_map_and_weight_map = _rasterize
    sys.modules["neural_renderer"] = nr
    sys.path.insert(0, REF)
    import utils.nmr as ref_nmr                        # the reference file, unmodified
    return ref_nmr


def main():
    torch.set_grad_enabled(False)
    ref_nmr = import_reference_nmr()
    R = ref_nmr.SMPLRenderer
    size, B = 64, 2                                    # (B = 3 would trip look_at's torch.cross default-dim quirk)
    v, f = S.uv_sphere()
    cam, verts = S.synthetic_frames(B + 1, seed=41, base_verts=v)
    tabs = S.synthetic_tables()
    src_img = S.synthetic_source(size)
    ns = types.SimpleNamespace(
        image_size=size, faces=f[None], map_fn=tabs["map_fn"], front_map_fn=tabs["front_map_fn"],
        back_map_fn=tabs["back_map_fn"], proj_func=ref_nmr.orthographic_proj_withz_idrot,
        eye=[0, 0, -(1. / np.tan(np.radians(30)) + 1)])
    ns.infer_face_index_map = lambda *a: (_ for _ in ()).throw(NotImplementedError())

    out = {}
    # source frame (frame 0): models/imitator.py:98-107
    s_f2v, s_fim, s_wim = R.render_fim_wim(ns, cam[:1], verts[:1].clone())
    s_cond, _ = R.encode_fim(ns, cam[:1], verts[:1], fim=s_fim, transpose=True)
    p2v = s_f2v[:, :, :, 0:2].clone()
    p2v[:, :, :, 1] *= -1
    vis = R.get_vis_f2pts(p2v, s_fim)
    # target frames: models/imitator.py:251-260
    f2v, fim, wim = R.render_fim_wim(ns, cam[1:], verts[1:].clone())
    cond, _ = R.encode_fim(ns, cam[1:], verts[1:], fim=fim, transpose=True)
    T = R.cal_bc_transform(ns, p2v.expand(B, -1, -1, -1), fim, wim)
    T_vis = R.cal_bc_transform(ns, vis.expand(B, -1, -1, -1), fim, wim)
    front = R.encode_front_fim(ns, fim, transpose=True, front_fn=True)
    back = R.encode_front_fim(ns, fim, transpose=True, front_fn=False)
    for ac in (True, False):
        tsf_img = F.grid_sample(src_img.expand(B, -1, -1, -1), T, align_corners=ac)
        out["tsf_inputs_ac%d" % int(ac)] = torch.cat([tsf_img, cond], dim=1).numpy()
    out.update(cam=cam.numpy(), src_fim=s_fim.numpy(), src_wim=s_wim.numpy(), src_cond=s_cond.numpy(),
               src_f2verts_sub=s_f2v[:, ::7].numpy(), vis_ids=(vis[0, :, 0, 0] != -2).nonzero()[:, 0].numpy().astype(np.int32),
               vis_sub=vis[:, ::7].numpy(), fim=fim.numpy(), wim=wim.numpy(), cond=cond.numpy(), T=T.numpy(),
               T_vis=T_vis.numpy(), front=front.numpy(), back=back.numpy(), f2verts_sub=f2v[:, ::7].numpy())

    # the restatement the CPU suite and the GPU box use must equal the reference's code on the full tensors
    o_f2v, o_fim, o_wim = nmr_ref.render_fim_wim(cam[1:], verts[1:], f, size)
    assert torch.equal(o_f2v, f2v) and torch.equal(o_fim, fim) and torch.equal(o_wim, wim)
    assert torch.equal(nmr_ref.encode_fim(o_fim, tabs["map_fn"]), cond)
    assert torch.equal(nmr_ref.cal_bc_transform(p2v.expand(B, -1, -1, -1), o_fim, o_wim, size), T)
    assert torch.equal(nmr_ref.get_vis_f2pts(p2v, s_fim), vis)
    assert torch.equal(nmr_ref.encode_fim(o_fim, tabs["front_map_fn"]), front)
    c = nmr_ref.correspond(cam[1:], verts[1:], f, tabs["map_fn"], p2v, src_img, size, align_corners=True)
    assert np.array_equal(c["tsf_inputs"].numpy(), out["tsf_inputs_ac1"])
    print("oracle/nmr_ref.py == reference utils/nmr.py methods on every tensor (exact)")
    np.savez_compressed(os.path.join(HERE, "nmr.npz"), **out)
    print("wrote nmr.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
