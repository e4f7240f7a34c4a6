"""Generates tests/golden/tasks.npz -- run ONLY in the build container (needs /root/reference).

Pins the three task classes -- motion imitation, novel-view synthesis and appearance transfer -- to the reference's OWN
code: ``models/imitator.py``, ``models/viewer.py`` and ``models/swapper.py`` are imported unmodified from /root/reference
and their methods run as unbound functions, on CPU, on a namespace that carries the attributes they read:

  Imitator.personalize / inference_by_smpls / transfer_params_by_smpl / swap_smpl / forward / warp_front
                       models/imitator.py:82-145, 192-268, 326-342 (cam strategies smooth / source / target, front_warp)

  Viewer.personalize   models/viewer.py:83-143      Swapper.personalize     models/swapper.py:99-165
  Viewer.rotate_trans  :237-244                     Swapper.swap            :199-239
  Viewer.view          :246-279                     Swapper.calculate_trans :242-253
  Viewer.forward       :281-288                     Swapper.forward / warp  :255-270
  Viewer.warp_front    :231-235

What stands in for the parts that cannot run here: ``.cuda()`` is the identity; ``self.render`` carries the reference's
``utils/nmr.py`` methods (unbound, as in make_nmr_golden.py) with the CPU rasterizer oracle behind
``nr.rasterize_face_index_map_and_weight_map`` (the reference's is CUDA-only); ``self.hmr`` is the exact-vertices
synthetic body model; ``self.generator`` is the reference's ``ImpersonatorGenerator`` on CPU with the synthetic weights;
the lookup tables are synthetic (the asset files are external downloads).  ``F.grid_sample`` without the flag means
align_corners=True as under the reference's pinned torch 1.2 (see make_generator_golden.py).

Images are 128 x 128 (the generator is fully convolutional): one golden run takes seconds.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
for m in ("ipdb", "h5py", "matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(m, types.ModuleType(m))

import make_nmr_golden as NG                                  # noqa: E402
from make_generator_golden import torch12_grid_sample         # noqa: E402  (also puts /root/reference on sys.path)
from impersonator_b200 import synthetic as S                  # noqa: E402

SIZE = 128
SRC_THETA = np.zeros(85, np.float32)
SRC_THETA[0:3] = (0.95, 0.03, -0.02)
TGT_THETA = np.zeros(85, np.float32)
TGT_THETA[0:3] = (1.05, -0.04, 0.05)
TGT_THETA[3] = 1.0                                            # a quarter turn: person B is seen from the side
IMIT_THETAS = np.zeros((3, 85), np.float32)                   # driving frames of the Imitator: quarter turns 0, 1, 3
IMIT_THETAS[:, 0] = (0.9, 1.0, 1.08)
IMIT_THETAS[:, 1] = (0.02, -0.05, 0.06)
IMIT_THETAS[:, 2] = (0.01, 0.04, -0.03)
IMIT_THETAS[:, 3] = (0.0, 1.0, 3.0)
VIEWS = [((10.0, 45.0, 10.0), (0.0, 0.0, 0.0)), ((0.0, 200.0, -5.0), (0.05, -0.02, 0.0))]       # degrees, translation


def sl(t):
    return t[:, :, 1::4, 2::4].contiguous().numpy()


def write_inputs(folder):
    a, b = os.path.join(folder, "a.png"), os.path.join(folder, "b.png")
    S.save_png(S.synthetic_source(160, seed=71)[0], a)                # another size: the loaders resize
    S.save_png(S.synthetic_source(SIZE, seed=72)[0], b)
    return a, b


def main():
    torch.set_grad_enabled(False)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    ref_nmr = NG.import_reference_nmr()
    from models.viewer import Viewer                          # the reference files, unmodified
    from models.swapper import Swapper
    from networks.generator import ImpersonatorGenerator
    R = ref_nmr.SMPLRenderer

    v, f = S.uv_sphere()
    tabs = S.synthetic_tables()
    gen = ImpersonatorGenerator(bg_dim=4, src_dim=6, tsf_dim=6, repeat_num=6).eval()
    gen.load_state_dict(S.fill_state_dict(gen.state_dict(), seed=0), strict=True)

    render = types.SimpleNamespace(
        image_size=SIZE, faces=f[None], map_fn=tabs["map_fn"], front_map_fn=tabs["front_map_fn"],
        back_map_fn=tabs["back_map_fn"], proj_func=ref_nmr.orthographic_proj_withz_idrot,
        eye=[0, 0, -(1. / np.tan(np.radians(30)) + 1)])
    for name in ("render_fim_wim", "encode_fim", "encode_front_fim", "cal_bc_transform"):
        setattr(render, name, types.MethodType(getattr(R, name), render))
    render.get_vis_f2pts = R.get_vis_f2pts
    render.infer_face_index_map = lambda *a: (_ for _ in ()).throw(NotImplementedError())

    def task(cls, **opt):
        ns = types.SimpleNamespace()
        ns._opt = types.SimpleNamespace(image_size=SIZE, bg_model="ORIGINAL", bg_ks=13, ft_ks=3, only_vis=False,
                                        front_warp=False, bg_replace=False)
        for k, val in opt.items():
            setattr(ns._opt, k, val)
        ns.hmr, ns.render, ns.detector, ns.generator, ns.bgnet = S.QuarterTurnBodyModel(v), render, None, gen, gen.bg_model
        for name in dir(cls):
            fn = getattr(cls, name)
            if callable(fn) and not name.startswith("__") and name not in ("personalize_",):
                try:
                    setattr(ns, name, types.MethodType(fn, ns))
                except TypeError:
                    pass
        return ns

    out = {}
    with tempfile.TemporaryDirectory() as tmp, torch12_grid_sample():
        a_png, b_png = write_inputs(tmp)

        # ---- motion imitation: the reference Imitator itself (models/imitator.py:82-145, 192-268, 326-342) ----
        from models.imitator import Imitator
        for tag, opt, strategy in (("smooth", {}, "smooth"), ("front_source", dict(front_warp=True), "source"),
                                   ("target", {}, "target"), ("only_vis", dict(only_vis=True), "smooth")):
            im = task(Imitator, **opt)
            im.src_info = im.tsf_info = im.first_cam = None
            im.personalize(a_png, src_smpl=SRC_THETA.copy())
            frames = im.inference_by_smpls([th.copy() for th in IMIT_THETAS], cam_strategy=strategy)
            assert len(frames) == 3 and frames[0].shape == (SIZE, SIZE, 3)
            for t, fr in enumerate(frames):
                out["imit_%s_%d" % (tag, t)] = fr[1::4, 2::4].copy()
            if tag == "smooth":
                out["imit_src_bg"] = sl(im.src_info["bg"])
                out["imit_last_T"] = im.tsf_info["T"][:, 1::4, 2::4].numpy()
                out["imit_last_cam"] = im.tsf_info["cam"].numpy()

        # ---- novel views ---------------------------------------------------------------------
        for tag, opt in (("plain", {}), ("front_bg", dict(front_warp=True, bg_replace=True))):
            vw = task(Viewer, **opt)
            vw.src_info = vw.tsf_info = vw.first_cam = None
            vw.personalize(a_png, src_smpl=SRC_THETA.copy())
            if tag == "plain":
                out["view_src_bg"] = sl(vw.src_info["bg"])
                out["view_src_cond"] = sl(vw.src_info["cond"])
            for i, (rt, t) in enumerate(VIEWS):
                rt = np.array(rt, dtype=np.float32) / 180 * np.pi          # run_view.py:34
                preds = vw.view(rt, np.array(t, dtype=np.float32), visualizer=None, name=str(i))
                out["view_%s_%d" % (tag, i)] = sl(preds)

        # ---- appearance transfer -------------------------------------------------------------
        part_info = S.synthetic_part_info()
        names = sorted(part_info.keys())
        part_fn = torch.zeros(f.shape[0] + 1, len(names) + 1)
        for i, name in enumerate(names):
            part_fn[part_info[name]["face"], i] = 1.0
        part_fn[-1, -1] = 1.0
        for tag, opt in (("plain", {}), ("front", dict(front_warp=True))):
            sw = task(Swapper, **opt)
            sw.PART_IDS = Swapper.PART_IDS
            sw.part_fn = part_fn
            sw.part_faces = [part_info[name]["face"] for name in names]
            sw.grid = R.create_meshgrid(SIZE)
            sw.src_info = sw.personalize(a_png, SRC_THETA.copy())
            sw.tsf_info = sw.personalize(b_png, TGT_THETA.copy())
            for part in ("body", "all"):
                preds = sw.swap(src_info=sw.src_info, tgt_info=sw.tsf_info, target_part=part, visualizer=None)
                out["swap_%s_%s" % (tag, part)] = sl(preds)
            if tag == "plain":
                mask = torch.sum(sw.src_info["part"][:, [0], ...], dim=1).bool()
                left = sorted(set(sw.part_faces[0]))
                T11, T21 = sw.calculate_trans(mask, left)
                out["swap_T11"] = T11[:, 1::4, 2::4].numpy()
                out["swap_T21"] = T21[:, 1::4, 2::4].numpy()
                out["swap_src_part"] = sw.src_info["part"][:, :, 1::4, 2::4].numpy()
                out["swap_tgt_bg"] = sl(sw.tsf_info["bg"])
    out["views"] = np.array(VIEWS, dtype=np.float32)
    out["src_theta"], out["tgt_theta"], out["imit_thetas"] = SRC_THETA, TGT_THETA, IMIT_THETAS
    np.savez_compressed(os.path.join(HERE, "tasks.npz"), **out)
    print("wrote tasks.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
