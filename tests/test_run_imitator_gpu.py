"""The reference's run_imitator.py main block (tests/run_imitator_body.py = lines 215-241 verbatim) executed against this
repo's ``Imitator``: ``Imitator(test_opt)`` builds generator (NetworksFactory + checkpoint), HMR (+SMPL) and the
SMPLRenderer (tables from the asset files) from ``opt`` like models/imitator.py:15-74; ``personalize(src_path)`` and
``inference(tgt_paths, tgt_smpls=None)`` run from image files through the HMR encoder.  All assets are synthetic files
in the real formats (impersonator_b200.synthetic.write_synthetic_assets).  The last frame is checked against the oracle."""
import glob
import os
import shutil
import types

import numpy as np
import pytest
import torch

from impersonator_b200 import synthetic as S
from run_imitator_body import BODY

pytestmark = pytest.mark.gpu


def reference_defaults(**kw):
    """options/base_options.py:12-60 + options/test_options.py:7-47 defaults (TestOptions().parse())."""
    d = dict(checkpoints_dir='./outputs/checkpoints/', map_name='uv_seg', part_info='assets/pretrains/smpl_part_info.json',
             uv_mapping='assets/pretrains/mapper.txt', hmr_model='assets/pretrains/hmr_tf2pt.pth',
             smpl_model='assets/pretrains/smpl_model.pkl', load_epoch=-1, load_path='', batch_size=4, time_step=10, tex_size=3,
             image_size=256, repeat_num=6, cond_nc=3, gpu_ids='0', model='imitator', name='running', gen_name='impersonator',
             norm_type='instance', is_train=False, output_dir='./outputs/results/', src_path='', tgt_path='', bg_model='ORIGINAL',
             bg_ks=13, ft_ks=3, only_vis=False, has_detector=False, body_seg=False, front_warp=False, post_tune=False,
             cam_strategy='smooth', ip='', port=31100, save_res=False)
    d.update(kw)
    return types.SimpleNamespace(**d)


def test_run_imitator_main_block(cuda, tmp_path, monkeypatch):
    torch.set_grad_enabled(False)
    from impersonator_b200.imitator import Imitator, morph
    A = S.write_synthetic_assets(str(tmp_path))
    monkeypatch.chdir(tmp_path)                                  # the reference resolves 'assets/pretrains/*' against the cwd
    opt = reference_defaults(src_path=A["src"], tgt_path=A["targets"], load_path=A["load_path"], save_res=True,
                             output_dir=str(tmp_path / "results"), batch_size=2)

    class TestOptions(object):
        def parse(self):
            return opt

    def mkdir(path):                                             # utils/util.py:215-226
        os.makedirs(path, exist_ok=True)
        return path

    def clear_dir(path):
        if os.path.exists(path):
            shutil.rmtree(path)
        return mkdir(path)

    def scan_tgt_paths(tgt_path, itv=20):                        # run_imitator.py:50-58
        if os.path.isdir(tgt_path):
            return sorted(glob.glob(os.path.join(tgt_path, '*')))[::itv]
        return [tgt_path]
    env = dict(TestOptions=TestOptions, Imitator=Imitator, mkdir=mkdir, clear_dir=clear_dir, scan_tgt_paths=scan_tgt_paths,
               os=os, VisdomVisualizer=None, adaptive_personalize=None)
    exec("if True:\n" + BODY, env)                               # the reference's main block, verbatim
    imitator = env["imitator"]

    saved = sorted(glob.glob(os.path.join(opt.output_dir, "imitators", "pred_*")))
    assert [os.path.basename(p) for p in saved] == ["pred_%03d.png" % i for i in range(3)]
    import cv2
    last = cv2.imread(saved[-1])
    assert last.shape == (256, 256, 3)

    # ---- oracle for the last frame (run_imitator.write_pair_info reads exactly these tsf_info / src_info entries)
    from impersonator_b200 import mesh
    from oracle import generator_ref as G, hmr_ref, nmr_ref, smpl_ref
    info, sinfo = imitator.tsf_info, imitator.src_info
    for k in ("fim", "T", "tsf_img", "theta", "j2d", "cam", "verts", "wim"):
        assert k in info, k
    assert info["fim"].shape == (1, 256, 256) and info["image"].shape == (320, 320, 3)

    def hmr_in(path):
        img = cv2.cvtColor(cv2.imread(path, -1), cv2.COLOR_BGR2RGB)
        return torch.from_numpy(cv2.resize(img, (224, 224)).astype(np.float32).transpose(2, 0, 1) / 255.0 * 2 - 1.0)[None]
    hsd = {k: v for k, v in A["hmr_state"].items() if not k.startswith("smpl.")}
    th_src = hmr_ref.forward(hmr_in(A["src"]), hsd)
    th_tgt = hmr_ref.forward(torch.cat([hmr_in(p) for p in A["target_files"]]), hsd)
    d_src = (sinfo["theta"].cpu() - th_src).abs().max().item()
    print("HMR theta of the source image vs oracle: %.3e" % d_src)
    assert d_src < 1e-3
    # geometry from here on uses the vertices the kernels produced (1e-4 theta differences move silhouette pixels)
    body = smpl_ref.model_tensors(S.synthetic_smpl_model(seed=3))
    cam = th_src[:, 0:3].clone()
    cam[:, 1:] += th_tgt[2:3, 1:3] - th_tgt[0:1, 1:3]            # swap_smpl 'smooth' (models/imitator.py:224-227)
    exp_theta = torch.cat([cam, th_tgt[2:3, 3:75], th_src[:, 75:]], dim=1)
    d_tsf = (info["theta"].cpu() - exp_theta).abs().max().item()
    print("tsf theta (HMR + smooth camera) vs oracle: %.3e" % d_tsf)
    assert d_tsf < 2e-3
    assert (smpl_ref.get_details(body, info["theta"].cpu())["verts"] - info["verts"].cpu()).abs().max() < 1e-5

    f = torch.from_numpy(np.load("assets/pretrains/smpl_faces.npy").astype(np.int32))
    map_fn = torch.from_numpy(mesh.create_mapping("uv_seg", "assets/pretrains/mapper.txt")).float()
    sd = A["generator_state"]
    src_img = sinfo["img"].cpu()
    f2v, sfim, _ = nmr_ref.render_fim_wim(sinfo["cam"].cpu(), sinfo["verts"].cpu(), f, 256)
    assert torch.equal(sfim, sinfo["fim"].cpu())
    cond = nmr_ref.encode_fim(sfim, map_fn)
    p2v = nmr_ref.src_p2verts(f2v)
    bg_mask = morph(cond[:, -1:], 13, 'erode')
    bg = G.resnet_generator(torch.cat([src_img * bg_mask, bg_mask], dim=1), sd, 'bg_model')
    ft_mask = 1 - morph(cond[:, -1:], 3, 'erode')
    feats = G.encode_src(torch.cat([src_img * ft_mask, cond], dim=1), sd)
    c = nmr_ref.correspond(info["cam"].cpu(), info["verts"].cpu(), f, map_fn, p2v, src_img, 256)
    assert torch.equal(c["fim"], info["fim"].cpu())
    cover = (c["fim"] >= 0).float().mean().item()
    assert cover > 0.01, "the synthetic body must be visible (%.4f)" % cover
    pred, _, _ = G.imitator_forward(bg, feats, c["tsf_inputs"], c["T"], sd)
    want = ((pred[0].permute(1, 2, 0).numpy()[..., ::-1] + 1) / 2.0 * 255).astype(np.uint8)       # save_cv2_img(normalize=True)
    diff = np.abs(last.astype(np.int32) - want.astype(np.int32))
    print("saved last frame vs oracle: max |du8| %d, pixels off by > 1: %d; coverage %.3f" % (diff.max(), int((diff > 1).sum()), cover))
    assert diff.max() <= 1                                       # 1e-3 on [-1,1] floats = 0.13 of a uint8 step (+ truncation)
