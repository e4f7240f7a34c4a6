"""oracle/tasks_ref.py (Viewer.view / Swapper.swap restated on CPU) against tests/golden/tasks.npz, the slices the
reference's own models/viewer.py and models/swapper.py produced on the same synthetic inputs."""
import numpy as np
import torch

import tasks_common as C
from impersonator_b200 import synthetic as S
from impersonator_b200.generator import ImpersonatorGenerator
from oracle import tasks_ref as T

# Network outputs: the restatement reproduces the reference's frames bit for bit in a quiet process (printed below), but
# oneDNN's fp32 convolutions are not run-to-run deterministic under load (observed: 5e-5 on the background net), so the
# gate is 2e-4 -- three orders of magnitude below what any orchestration error produces.  Flows / tables stay exact.
NET_TOL = 2e-4


def _setup(tmp_path):
    torch.set_grad_enabled(False)
    g = np.load(C.GOLD)
    v, f = S.uv_sphere()
    tabs = S.synthetic_tables()
    sd = S.fill_state_dict(ImpersonatorGenerator(bg_dim=4, src_dim=6, tsf_dim=6, repeat_num=6).state_dict(), seed=0)
    a_png, b_png = C.write_inputs(tmp_path)
    body = S.QuarterTurnBodyModel(v)
    return g, v, f, tabs, sd, a_png, b_png, body


def test_oracle_view_matches_reference_viewer(tmp_path):
    g, v, f, tabs, sd, a_png, _, body = _setup(tmp_path)
    d = body.get_details(torch.from_numpy(g["src_theta"])[None])
    info = T.personalize(C.read_like_reference(a_png), d["cam"], d["verts"], f, tabs, sd, C.SIZE, "viewer")
    assert np.abs(C.sl(info["bg"]) - g["view_src_bg"]).max() < NET_TOL
    assert np.array_equal(C.sl(info["cond"]), g["view_src_cond"])
    for tag, kw in (("plain", {}), ("front_bg", dict(front_warp=True, bg_replace=True))):
        for i, (rt, t) in enumerate(g["views"]):
            preds, _ = T.view(info, rt / 180 * np.pi, t, f, tabs, sd, C.SIZE, **kw)
            err = np.abs(C.sl(preds) - g["view_%s_%d" % (tag, i)]).max()
            print("view %s %d: oracle vs reference max-abs %.2e" % (tag, i, err))
            assert err < NET_TOL


def test_oracle_swap_matches_reference_swapper(tmp_path):
    g, v, f, tabs, sd, a_png, b_png, body = _setup(tmp_path)
    _, part_fn, part_faces = C.part_table(f.shape[0])
    infos = []
    for png, key in ((a_png, "src_theta"), (b_png, "tgt_theta")):
        d = body.get_details(torch.from_numpy(g[key])[None])
        infos.append(T.personalize(C.read_like_reference(png), d["cam"], d["verts"], f, tabs, sd, C.SIZE, "swapper", part_fn))
    src, tgt = infos
    assert np.array_equal(src["part"][:, :, 1::4, 2::4].numpy(), g["swap_src_part"])
    assert np.abs(C.sl(tgt["bg"]) - g["swap_tgt_bg"]).max() < NET_TOL
    for tag, fw in (("plain", False), ("front", True)):
        for part in ("body", "all"):
            preds, T11, T21 = T.swap(src, tgt, part_faces, tabs, sd, C.SIZE, part, front_warp=fw)
            err = np.abs(C.sl(preds) - g["swap_%s_%s" % (tag, part)]).max()
            print("swap %s %s: oracle vs reference max-abs %.2e" % (tag, part, err))
            assert err < NET_TOL
            if tag == "plain" and part == "body":
                assert np.array_equal(T11[:, 1::4, 2::4].numpy(), g["swap_T11"])
                assert np.abs(T21[:, 1::4, 2::4].numpy() - g["swap_T21"]).max() < 1e-6


def test_product_host_helpers_match_oracle():
    """Device-agnostic host logic of the mirrors (no kernels involved): Euler matrix, rigid transform, camera swap."""
    from impersonator_b200.swapper import Swapper
    from impersonator_b200.viewer import Viewer, euler2matrix
    rs = np.random.RandomState(3)
    for _ in range(5):
        rt = (rs.rand(3) * 2 - 1) * np.pi
        assert np.array_equal(euler2matrix(rt.astype(np.float32)), T.euler2matrix(rt.astype(np.float32)))
    X = torch.from_numpy(rs.randn(1, 50, 3).astype(np.float32))
    rt, t = np.array([0.2, -1.1, 0.4], np.float32), np.array([0.1, 0.0, -0.3], np.float32)
    got = Viewer.rotate_trans(None, rt, t, X)
    want = torch.bmm(X, torch.from_numpy(T.euler2matrix(rt))[None]) + torch.from_numpy(t)[None, None]
    assert (got - want).abs().max().item() < 1e-6
    many = Viewer.rotate_trans(None, np.stack([rt, rt * 2]), t, X)
    assert many.shape == (2, 50, 3) and torch.equal(many[0:1], got)
    src_cam = torch.tensor([[0.9, 0.1, -0.2]])
    tgt = torch.zeros(1, 85)
    tgt[0, 0:3] = torch.tensor([1.2, 0.05, 0.02])
    keep = tgt.clone()
    out = Swapper.swap_smpl(None, src_cam, torch.zeros(1, 10), tgt)
    assert torch.equal(tgt, keep)                                       # the caller's vector is not edited
    assert torch.allclose(out[0, 0:3], torch.tensor([0.9, 0.15, -0.18]))   # models/swapper.py:183-186: scale of A, shifts add up


def test_oracle_imitation_matches_reference_imitator(tmp_path):
    """The oracle's per-frame loop against what the reference's own Imitator.personalize + inference_by_smpls returned
    (camera strategies smooth / source / target, with and without front_warp)."""
    g, v, f, tabs, sd, a_png, _, body = _setup(tmp_path)
    d = body.get_details(torch.from_numpy(g["src_theta"])[None])
    info = T.personalize(C.read_like_reference(a_png), d["cam"], d["verts"], f, tabs, sd, C.SIZE, "imitator")
    assert np.abs(C.sl(info["bg"]) - g["imit_src_bg"]).max() < NET_TOL
    thetas = torch.from_numpy(g["imit_thetas"])
    vis = T.personalize(C.read_like_reference(a_png), d["cam"], d["verts"], f, tabs, sd, C.SIZE, "imitator", only_vis=True)
    for tag, strategy, fw in (("smooth", "smooth", False), ("front_source", "source", True), ("target", "target", False),
                              ("only_vis", "smooth", False)):
        outs, lastT = T.imitate(vis if tag == "only_vis" else info, d["shape"], thetas, body, f, tabs, sd, C.SIZE, strategy,
                                front_warp=fw)
        for t, fr in enumerate(outs):
            err = np.abs(fr[1::4, 2::4] - g["imit_%s_%d" % (tag, t)]).max()
            assert err < NET_TOL, (tag, t, err)
        if tag == "smooth":
            assert np.abs(lastT[:, 1::4, 2::4].numpy() - g["imit_last_T"]).max() < 1e-6


def test_viewer_and_swapper_host_logic_on_cpu(tmp_path, monkeypatch):
    """The product's Viewer / Swapper classes themselves, on CPU: every kernel front-end replaced by its torch stand-in
    (tests/kernel_emulator.py, contracts of include/lwb_b200.h), compared with what the reference's models/viewer.py and
    models/swapper.py produced (tests/golden/tasks.npz).  Covers the orchestration the GPU tests cover, without a GPU:
    personalize (masks, background paste, part map), rotate_trans, bg_replace / front_warp, T11 / T21, the two-source swap."""
    import kernel_emulator
    from impersonator_b200.nmr import SMPLRenderer
    from impersonator_b200.swapper import Swapper
    from impersonator_b200.viewer import Viewer
    kernel_emulator.install_tasks(monkeypatch)
    g, v, f, tabs, sd, a_png, b_png, body = _setup(tmp_path)
    net = ImpersonatorGenerator(bg_dim=4, src_dim=6, tsf_dim=6, repeat_num=6).eval()
    net.load_state_dict(sd)

    def render(front):
        return SMPLRenderer(image_size=C.SIZE, faces=f.numpy(), map_fn=tabs["map_fn"], has_front=front,
                            front_map_fn=tabs["front_map_fn"], back_map_fn=tabs["back_map_fn"])
    worst = 0.0
    for tag, front, bg_replace in (("plain", False, False), ("front_bg", True, True)):
        opt = C.Opt()
        opt.front_warp, opt.bg_replace = front, bg_replace
        vw = Viewer(opt, generator=net, hmr=body, render=render(front), device="cpu")
        vw.personalize(a_png, src_smpl=g["src_theta"].copy())
        assert np.array_equal(C.sl(vw.src_info["cond"]), g["view_src_cond"])
        for i, (rt, t) in enumerate(g["views"]):
            preds = vw.view(rt / 180 * np.pi, t, name=str(i))
            worst = max(worst, float(np.abs(C.sl(preds) - g["view_%s_%d" % (tag, i)]).max()))
    part_info, _, part_faces = C.part_table(f.shape[0])
    for tag, front in (("plain", False), ("front", True)):
        opt = C.Opt()
        opt.front_warp = front
        sw = Swapper(opt, part_info=part_info, generator=net, hmr=body, render=render(front), device="cpu")
        sw.swap_setup(a_png, b_png, src_smpl=g["src_theta"].copy(), tgt_smpl=g["tgt_theta"].copy())
        assert np.array_equal(sw.src_info["part"][:, :, 1::4, 2::4].numpy(), g["swap_src_part"])
        for part in ("body", "all"):
            preds = sw.swap(sw.src_info, sw.tsf_info, target_part=part)
            worst = max(worst, float(np.abs(C.sl(preds) - g["swap_%s_%s" % (tag, part)]).max()))
        if tag == "plain":
            mask = torch.sum(sw.src_info["part"][:, [0], ...], dim=1).bool()
            T11, T21 = sw.calculate_trans(mask, sorted(set(part_faces[0])))
            assert np.array_equal(T11[:, 1::4, 2::4].numpy(), g["swap_T11"])
            assert np.abs(T21[:, 1::4, 2::4].numpy() - g["swap_T21"]).max() < 1e-6
    print("Viewer / Swapper on the kernel emulator vs the reference's frames: max-abs %.2e" % worst)
    assert worst < 2e-4


def test_imitator_per_frame_api_host_logic_on_cpu(tmp_path, monkeypatch):
    """The Imitator mirror's per-frame methods as the reference's loop calls them (models/imitator.py:198-203:
    transfer_params_by_smpl + forward), on the kernel emulator, against the reference Imitator's frames.  (The chunked
    ``inference`` drives CUDA streams / pinned memory and is covered by the GPU tests.)"""
    import kernel_emulator
    from impersonator_b200.imitator import Imitator
    from impersonator_b200.nmr import SMPLRenderer
    kernel_emulator.install_tasks(monkeypatch)
    g, v, f, tabs, sd, a_png, _, body = _setup(tmp_path)
    net = ImpersonatorGenerator(bg_dim=4, src_dim=6, tsf_dim=6, repeat_num=6).eval()
    net.load_state_dict(sd)
    worst = 0.0
    for tag, strategy, front in (("smooth", "smooth", False), ("front_source", "source", True), ("target", "target", False),
                                 ("only_vis", "smooth", False)):
        opt = C.Opt()
        opt.front_warp, opt.only_vis = front, tag == "only_vis"
        render = SMPLRenderer(image_size=C.SIZE, faces=f.numpy(), map_fn=tabs["map_fn"], has_front=front,
                              front_map_fn=tabs["front_map_fn"], back_map_fn=tabs["back_map_fn"])
        im = Imitator(opt, generator=net, hmr=body, render=render, device="cpu")
        im.personalize(a_png, src_smpl=g["src_theta"].copy())
        for t, th in enumerate(g["imit_thetas"]):
            tsf_inputs = im.transfer_params_by_smpl(th.copy(), strategy, t=t)
            preds = im.forward(tsf_inputs, im.tsf_info['T'])
            fr = preds[0].permute(1, 2, 0).numpy()
            worst = max(worst, float(np.abs(fr[1::4, 2::4] - g["imit_%s_%d" % (tag, t)]).max()))
        if tag == "smooth":
            assert np.abs(im.tsf_info["T"][:, 1::4, 2::4].numpy() - g["imit_last_T"]).max() < 1e-6
            # demo_view.py:55-67 drives the Imitator with float64 [cam | pose] vectors of 75 entries (no shape part)
            short = im.transfer_params_by_smpl(g["imit_thetas"][2][:75].astype(np.float64), strategy, t=2)
            assert torch.equal(short, tsf_inputs)
    print("Imitator per-frame API on the kernel emulator vs the reference's frames: max-abs %.2e" % worst)
    assert worst < 2e-4
