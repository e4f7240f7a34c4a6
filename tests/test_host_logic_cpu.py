"""Host-side logic of the Imitator mirror that needs no GPU: camera strategies of swap_smpl (models/imitator.py:216-234),
the erode/dilate of utils/util.py:73-89, frame chunking, and the synthetic SMPL model file format."""
import numpy as np
import torch

from impersonator_b200.imitator import Imitator, morph
from impersonator_b200 import synthetic as S


def ref_swap_smpl(first_cam, src_cam, src_shape, tgt_smpl, cam_strategy):      # models/imitator.py:216-234, one frame
    tgt_cam = tgt_smpl[:, 0:3].contiguous()
    pose = tgt_smpl[:, 3:75].contiguous()
    if cam_strategy == 'smooth':
        cam = src_cam.clone()
        cam[:, 1:] += tgt_cam[:, 1:] - first_cam[:, 1:]
    elif cam_strategy == 'source':
        cam = src_cam
    else:
        cam = tgt_cam
    return torch.cat([cam, pose, src_shape], dim=1)


def test_swap_smpl_batched_equals_reference_per_frame():
    g = torch.Generator().manual_seed(0)
    im = Imitator.__new__(Imitator)
    im.first_cam = torch.rand(1, 3, generator=g)
    src_cam, src_shape = torch.rand(1, 3, generator=g), torch.rand(1, 10, generator=g)
    tgt = torch.rand(5, 85, generator=g)
    for strategy in ('smooth', 'source', 'copy'):
        got = im.swap_smpl(src_cam, src_shape, tgt, cam_strategy=strategy)
        ref = torch.cat([ref_swap_smpl(im.first_cam, src_cam, src_shape, tgt[i:i + 1], strategy) for i in range(5)])
        assert got.shape == (5, 85) and torch.equal(got, ref)


def test_morph_equals_reference_box_filter():
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(1)
    mask = (torch.rand(2, 1, 40, 40, generator=g) > 0.3).float()
    for ks in (3, 13):
        pad = ks // 2
        kernel = torch.ones(1, 1, ks, ks)
        erode = (F.conv2d(F.pad(mask, [pad] * 4, value=1.0), kernel) == ks * ks).float()      # utils/util.py:76-82
        dilate = (F.conv2d(F.pad(mask, [pad] * 4, value=0.0), kernel) >= 1).float()           # utils/util.py:84-89
        assert torch.equal(morph(mask, ks, 'erode'), erode)
        assert torch.equal(morph(mask, ks, 'dilate'), dilate)


def test_chunks_cover_the_sequence_in_order():
    im = Imitator.__new__(Imitator)

    class Opt(object):
        batch_size = 16
    im._opt = Opt()
    for n in (0, 1, 16, 17, 50):
        ch = im._chunks(n)
        assert [i for a, b in ch for i in range(a, b)] == list(range(n))
        assert all(b - a <= 16 for a, b in ch)


def test_synthetic_smpl_model_has_the_pickle_layout():
    """Same keys / shapes / dtypes that SMPL.__init__ reads from smpl_model.pkl (networks/batch_smpl.py:236-283)."""
    dd = S.synthetic_smpl_model(seed=3)
    assert dd['v_template'].shape == (6890, 3) and dd['shapedirs'].shape == (6890, 3, 10)
    assert dd['posedirs'].shape == (6890, 3, 207) and dd['weights'].shape == (6890, 24)
    assert dd['J_regressor'].shape == (24, 6890) and dd['cocoplus_regressor'].shape == (19, 6890)
    assert dd['kintree_table'].shape == (2, 24) and dd['kintree_table'][0, 0] == 4294967295
    assert np.allclose(dd['weights'].sum(1), 1) and np.allclose(np.asarray(dd['J_regressor'].sum(1)).ravel(), 1)
    assert dd['f'].shape == (13776, 3)
    th = S.synthetic_smpl_params(3, seed=1)
    assert th.shape == (3, 85) and th.dtype == torch.float32


def test_oracle_self_correspondence_is_the_identity_warp():
    """Domain round trip (size independent): a frame corresponded with ITSELF must give T = pixel-centre coordinates on
    every covered pixel (cal_bc_transform, utils/nmr.py:617-659, inverts the rasterizer's barycentrics) and therefore
    warp the source image onto itself; uncovered pixels keep -2.  Pins the oracle's conventions (y flip, row flip)
    without any golden file.  The image identity needs sampling at pixel centres (align_corners=False); the reference's
    torch-1.2 convention (the default elsewhere) samples up to half a pixel off, as it did upstream."""
    from oracle import nmr_ref
    size = 128
    v, f = S.uv_sphere()
    cam, verts = S.synthetic_frames(1, seed=8, base_verts=v)
    tabs = S.synthetic_tables()
    ys, xs = torch.meshgrid(torch.arange(size, dtype=torch.float32), torch.arange(size, dtype=torch.float32), indexing="ij")
    gx, gy = (2 * xs + 1 - size) / size, (2 * ys + 1 - size) / size
    src = torch.stack([torch.sin(3 * gx) * torch.cos(2 * gy), gx * gy, torch.cos(4 * gx + gy)])[None]
    f2v, fim, _ = nmr_ref.render_fim_wim(cam, verts, f, size)
    out = nmr_ref.correspond(cam, verts, f, tabs["map_fn"], nmr_ref.src_p2verts(f2v), src, size, align_corners=False)
    cov = fim[0] >= 0
    assert 0.05 < cov.float().mean() < 0.6
    T = out["T"][0]
    assert (T[..., 0] - gx)[cov].abs().max() < 5e-4 and (T[..., 1] - gy)[cov].abs().max() < 5e-4
    assert torch.all(T[~cov] == -2)
    assert (out["tsf_img"][0] - src[0])[:, cov].abs().max() < 2e-3
    assert torch.all(out["tsf_img"][0][:, ~cov] == 0)


def test_weight_exponent_rule():
    """Per-layer scale of the fp16f8 weight packing: max|w| * 2^E in [2^14, 2^15) for any magnitude."""
    from impersonator_b200 import kernels as K
    for a, e in ((1.5, 14), (0.99, 15), (0.5, 15), (0.02, 20), (2.0, 13), (5.0, 12), (0.0, 15)):
        assert K.weight_exponent(a) == e, (a, K.weight_exponent(a))
        if a > 0:
            assert 2 ** 14 <= a * 2.0 ** e < 2 ** 15
