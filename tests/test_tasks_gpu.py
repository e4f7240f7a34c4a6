"""GPU parity of the Viewer and Swapper mirrors (impersonator_b200/viewer.py, swapper.py) against
  * tests/golden/tasks.npz -- what the reference's own models/viewer.py / models/swapper.py produce on these inputs, and
  * oracle/tasks_ref.py on the full tensors.
Bar: 1e-3 max-abs on pixels (BASELINE.json north_star); flows 1e-5."""
import numpy as np
import pytest
import torch

import tasks_common as C
from impersonator_b200 import synthetic as S
from impersonator_b200.generator import ImpersonatorGenerator
from impersonator_b200.nmr import SMPLRenderer
from impersonator_b200.swapper import Swapper
from impersonator_b200.viewer import Viewer
from oracle import tasks_ref as T

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module")
def world(cuda, tmp_path_factory):
    torch.set_grad_enabled(False)
    v, f = S.uv_sphere()
    tabs = S.synthetic_tables()
    net = ImpersonatorGenerator(bg_dim=4, src_dim=6, tsf_dim=6, repeat_num=6)
    sd = S.fill_state_dict(net.state_dict(), seed=0)
    net.load_state_dict(sd)
    a_png, b_png = C.write_inputs(tmp_path_factory.mktemp("tasks"))
    return dict(v=v, f=f, tabs=tabs, net=net.to(cuda).eval(), sd=sd, a=a_png, b=b_png, g=np.load(C.GOLD))


def _render(w, front):
    return SMPLRenderer(image_size=C.SIZE, faces=w["f"].numpy(), map_fn=w["tabs"]["map_fn"], has_front=front,
                        front_map_fn=w["tabs"]["front_map_fn"], back_map_fn=w["tabs"]["back_map_fn"])


@pytest.mark.parametrize("tag,front,bg_replace", [("plain", False, False), ("front_bg", True, True)])
def test_viewer_views_match_reference_and_oracle(cuda, world, tag, front, bg_replace):
    w, g = world, world["g"]
    opt = C.Opt()
    opt.front_warp, opt.bg_replace = front, bg_replace
    vw = Viewer(opt, generator=w["net"], hmr=S.QuarterTurnBodyModel(w["v"]), render=_render(w, front), device=cuda)
    vw.personalize(w["a"], src_smpl=g["src_theta"].copy())
    assert np.abs(C.sl(vw.src_info["bg"]) - g["view_src_bg"]).max() < TOL
    assert np.array_equal(C.sl(vw.src_info["cond"]), g["view_src_cond"])
    d = S.QuarterTurnBodyModel(w["v"]).get_details(torch.from_numpy(g["src_theta"])[None])
    info = T.personalize(C.read_like_reference(w["a"]), d["cam"], d["verts"], w["f"], w["tabs"], w["sd"], C.SIZE, "viewer")
    singles = []
    for i, (rt, t) in enumerate(g["views"]):
        rad = rt / 180 * np.pi
        preds = vw.view(rad, t, name=str(i))
        singles.append(preds.clone())
        e_gold = np.abs(C.sl(preds) - g["view_%s_%d" % (tag, i)]).max()
        # the oracle consumes the vertices the product's own rotate_trans produced: a 1e-7 difference between the CPU and
        # GPU matmul could flip a silhouette pixel of the bit-exact rasterizer
        mesh = vw.tsf_info["verts"].cpu()
        c = T.nmr_ref.correspond(info["cam"], mesh, w["f"], w["tabs"]["map_fn"], info["p2verts"], info["img"], C.SIZE)
        bg = info["bg"] if bg_replace else torch.zeros_like(info["bg"])
        ref, _, mask = T.G.imitator_forward(bg, info["feats"], c["tsf_inputs"], c["T"], w["sd"])
        if front:
            fm = T.nmr_ref.encode_fim(c["fim"], w["tabs"]["front_map_fn"])
            ref = (1 - fm) * ref + c["tsf_img"] * fm * (1 - mask)
        e_or = (preds.cpu() - ref).abs().max().item()
        print("view %s %d: vs reference golden %.2e, vs oracle (full) %.2e" % (tag, i, e_gold, e_or))
        assert e_or < TOL
        assert e_gold < TOL
        assert torch.equal(vw.tsf_info["fim"].cpu(), c["fim"])
    # the views of run_view.py's loop as ONE batch
    rts = np.stack([rt / 180 * np.pi for rt, _ in g["views"]])
    ts = np.stack([t for _, t in g["views"]])
    both = vw.view_many(rts, ts)
    assert both.shape[0] == 2
    for i in range(2):
        assert (both[i:i + 1] - singles[i]).abs().max().item() < 1e-5


@pytest.mark.parametrize("tag,front", [("plain", False), ("front", True)])
def test_swapper_matches_reference_and_oracle(cuda, world, tag, front):
    w, g = world, world["g"]
    opt = C.Opt()
    opt.front_warp = front
    part_info, part_fn, part_faces = C.part_table(w["f"].shape[0])
    sw = Swapper(opt, part_info=part_info, generator=w["net"], hmr=S.QuarterTurnBodyModel(w["v"]),
                 render=_render(w, front), device=cuda)
    sw.swap_setup(w["a"], w["b"], src_smpl=g["src_theta"].copy(), tgt_smpl=g["tgt_theta"].copy())
    assert np.array_equal(sw.src_info["part"][:, :, 1::4, 2::4].cpu().numpy(), g["swap_src_part"])
    assert np.abs(C.sl(sw.tsf_info["bg"]) - g["swap_tgt_bg"]).max() < TOL
    body = S.QuarterTurnBodyModel(w["v"])
    infos = []
    for png, key in ((w["a"], "src_theta"), (w["b"], "tgt_theta")):
        d = body.get_details(torch.from_numpy(g[key])[None])
        infos.append(T.personalize(C.read_like_reference(png), d["cam"], d["verts"], w["f"], w["tabs"], w["sd"], C.SIZE,
                                   "swapper", part_fn))
    for part in ("body", "all"):
        preds = sw.swap(sw.src_info, sw.tsf_info, target_part=part)
        ref, T11, T21 = T.swap(infos[0], infos[1], part_faces, w["tabs"], w["sd"], C.SIZE, part, front_warp=front)
        e_gold = np.abs(C.sl(preds) - g["swap_%s_%s" % (tag, part)]).max()
        e_or = (preds.cpu() - ref).abs().max().item()
        e_T = max((sw.T11.cpu() - T11).abs().max().item(), (sw.T21.cpu() - T21).abs().max().item())
        print("swap %s %s: vs reference golden %.2e, vs oracle (full) %.2e, flows %.1e" % (tag, part, e_gold, e_or, e_T))
        assert e_T < 1e-5 and e_or < TOL and e_gold < TOL
    if tag == "plain":
        mask = torch.sum(sw.src_info["part"][:, [0], ...], dim=1).bool()
        T11, T21 = sw.calculate_trans(mask, sorted(set(part_faces[0])))
        assert np.array_equal(T11[:, 1::4, 2::4].cpu().numpy(), g["swap_T11"])
        assert np.abs(T21[:, 1::4, 2::4].cpu().numpy() - g["swap_T21"]).max() < 1e-5


def test_run_view_and_run_swap_from_asset_files(cuda, tmp_path, monkeypatch):
    """``Viewer(opt)`` / ``Swapper(opt)`` built from files exactly like the reference (models/viewer.py:27-76,
    models/swapper.py:22-89: checkpoint, HMR + SMPL pickle, renderer tables, part table from ``opt.uv_mapping``), then the
    loops of run_view.py:52-73 (16 views, 360 / 16 degrees apart) and run_swap.py:52-63, SMPL estimated from the images."""
    from test_run_imitator_gpu import reference_defaults
    torch.set_grad_enabled(False)
    A = S.write_synthetic_assets(str(tmp_path))
    monkeypatch.chdir(tmp_path)
    opt = reference_defaults(src_path=A["src"], tgt_path=A["target_files"][0], load_path=A["load_path"], bg_replace=True,
                             swap_part='body', view_params='R=0,90,0/t=0,0,0')
    viewer = Viewer(opt=opt)
    viewer.personalize(opt.src_path, visualizer=None)
    length, delta = 16, 360 / 16
    R = np.zeros((length, 3), np.float32)
    R[:, 0] = R[:, 2] = 10 / 180 * np.pi
    R[:, 1] = delta * np.arange(length) / 180.0 * np.pi
    t = np.zeros(3, np.float32)
    pred_outs = torch.cat([viewer.view(R[i], t, visualizer=None, name=str(i)) for i in range(length)], dim=0)
    assert pred_outs.shape == (16, 3, 256, 256) and torch.isfinite(pred_outs).all()
    batched = viewer.view_many(R, t)                                   # the same 16 views as one batch of 16
    assert (batched - pred_outs).abs().max().item() < 1e-4
    assert (pred_outs[0] - pred_outs[8]).abs().max().item() > 0.05     # front and back views differ

    swapper = Swapper(opt=opt)
    assert swapper.part_fn.shape == (S.SMPL_F + 1, 11) and len(swapper.part_faces) == 10
    swapper.swap_setup(opt.src_path, opt.tgt_path)
    preds = swapper.swap(src_info=swapper.src_info, tgt_info=swapper.tsf_info, target_part=opt.swap_part, visualizer=None)
    assert preds.shape == (1, 3, 256, 256) and torch.isfinite(preds).all() and preds.abs().max().item() <= 1.5
    other = swapper.swap(src_info=swapper.src_info, tgt_info=swapper.tsf_info, target_part='all')
    assert (other - preds).abs().max().item() > 1e-3                   # part 0 comes from person A in 'body' mode only


@pytest.mark.parametrize("tag,strategy,front,batch", [("smooth", "smooth", False, 4), ("front_source", "source", True, 2),
                                                      ("target", "target", False, 1)])
def test_imitator_matches_reference_imitator(cuda, world, tag, strategy, front, batch):
    """``Imitator.personalize`` + ``inference_by_smpls`` against the frames the reference's own ``models/imitator.py`` returned
    for the same source image file, SMPL vectors and weights (tests/golden/tasks.npz): three camera strategies, with and
    without front_warp, chunk sizes 4 / 2 / 1 (the reference runs frame by frame)."""
    from impersonator_b200.imitator import Imitator
    w, g = world, world["g"]
    opt = C.Opt()
    opt.front_warp, opt.batch_size = front, batch
    im = Imitator(opt, generator=w["net"], hmr=S.QuarterTurnBodyModel(w["v"]), render=_render(w, front), device=cuda)
    im.personalize(w["a"], src_smpl=g["src_theta"].copy())
    assert np.abs(C.sl(im.src_info["bg"]) - g["imit_src_bg"]).max() < TOL
    frames = im.inference_by_smpls([th.copy() for th in g["imit_thetas"]], cam_strategy=strategy)
    assert len(frames) == 3
    errs = [float(np.abs(fr[1::4, 2::4] - g["imit_%s_%d" % (tag, t)]).max()) for t, fr in enumerate(frames)]
    print("Imitator %s (chunks of %d) vs the reference Imitator: %s" % (tag, batch, ["%.2e" % e for e in errs]))
    assert max(errs) < TOL
    if tag == "smooth":
        assert np.abs(im.tsf_info["T"][:, 1::4, 2::4].cpu().numpy() - g["imit_last_T"]).max() < 1e-5
        assert np.abs(im.tsf_info["cam"].cpu().numpy() - g["imit_last_cam"]).max() < 1e-6
