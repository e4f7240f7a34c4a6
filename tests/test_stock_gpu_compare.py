"""the "kernel to beat" on the same B200 -- the reference's own rasterizer kernels (oracle/_ref,
compiled unmodified for sm_100a) and the reference generator's ATen ops on the GPU (cuDNN; strict fp32 and TF32) --
timed next to this library on the BASELINE config-3 batch (16 frames, 256x256).  Writes gpurun_out/stock_compare.json.
The stock path is the checker/baseline only; nothing here is shipped."""
import json
import os

import pytest
import torch

from impersonator_b200 import synthetic as S
from impersonator_b200.generator import ImpersonatorGenerator
from impersonator_b200.nmr import SMPLRenderer
from oracle import generator_ref as G
from oracle import nmr_ref, raster

pytestmark = pytest.mark.gpu
FULL = os.environ.get("LWB_RUN_STOCK") == "1"      # also the TF32 leg and the reference rasterizer kernels (~2 min more)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def test_stock_torch_and_reference_kernels_vs_this_library(cuda):
    torch.set_grad_enabled(False)
    B, size = 16, 256
    n = ImpersonatorGenerator(bg_dim=4, src_dim=6, tsf_dim=6, repeat_num=6)
    sd = S.fill_state_dict(n.state_dict(), seed=0)
    n.load_state_dict(sd)
    n = n.to(cuda).eval()
    sd_gpu = {k: v.to(cuda) for k, v in sd.items()}
    inp = S.synthetic_generator_inputs(B, size, seed=21)
    src, tsf, T = inp["src"][:1].to(cuda), inp["tsf"].to(cuda), inp["T"].to(cuda)
    out = {"batch": B, "image_size": size}

    # --- generator.inference: stock ATen/cuDNN (the reference modules' ops) -------------------------------
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    enc_o, res_o = G.encode_src(src, sd_gpu)
    enc_b = [e.expand(B, -1, -1, -1).contiguous() for e in enc_o]          # reference grid_sample needs equal N
    res_b = [r.expand(B, -1, -1, -1).contiguous() for r in res_o]
    torch.backends.cudnn.allow_tf32 = False                     # strict fp32 first: this run is also the parity reference
    torch.backends.cuda.matmul.allow_tf32 = False
    ref_img, ref_mask = G.inference(enc_b, res_b, tsf, T, sd_gpu)
    out["stock_generator_ms_fp32"] = timeit(lambda: G.inference(enc_b, res_b, tsf, T, sd_gpu), iters=3, warm=1)
    if FULL:
        torch.backends.cudnn.allow_tf32 = True                  # torch's default on Ampere+
        torch.backends.cuda.matmul.allow_tf32 = True
        tf_img, _ = G.inference(enc_b, res_b, tsf, T, sd_gpu)
        out["stock_generator_ms_tf32"] = timeit(lambda: G.inference(enc_b, res_b, tsf, T, sd_gpu))
        out["stock_tf32_vs_stock_fp32_max_abs"] = (tf_img - ref_img).abs().max().item()
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False

    enc, res = n.encode_src(src)
    out["lwb_generator_ms"] = timeit(lambda: n.inference(enc, res, tsf, T))
    img, mask = n.inference(enc, res, tsf, T)
    per_frame = torch.maximum((img - ref_img).abs().amax(dim=(1, 2, 3)), (mask - ref_mask).abs().amax(dim=(1, 2, 3)))
    out["lwb_vs_stock_fp32_max_abs"] = per_frame.max().item()
    out["lwb_vs_stock_fp32_per_frame"] = [round(v, 6) for v in per_frame.tolist()]
    assert out["lwb_vs_stock_fp32_max_abs"] < 1e-3               # all 16 frames of the headline batch, every pixel

    # --- rasterizer: the reference's kernels vs k_face_raster ------------------------------------------------
    if FULL and raster.gpu_ref_available():
        v, f = S.uv_sphere()
        cam, verts = S.synthetic_frames(B, seed=1234, base_verts=v)
        faces = nmr_ref.project_to_faces(cam, verts, f).to(cuda).contiguous()
        out["reference_raster_kernels_ms"] = timeit(lambda: raster.forward_face_index_map_gpu_ref(faces, size), iters=5, warm=2)
        tabs = S.synthetic_tables()
        r = SMPLRenderer(image_size=size, faces=f, map_fn=tabs["map_fn"]).to(cuda)
        cam_d, verts_d = cam.to(cuda), verts.to(cuda)
        out["lwb_render_fim_wim_ms"] = timeit(lambda: r.render_fim_wim(cam_d, verts_d))
    # --- InpaintSANet (once per source): the reference module's ATen ops (cuDNN / cuBLAS) vs the conv-engine stream ----------
    from impersonator_b200.inpaintor import InpaintSANet
    from oracle import inpaintor_ref as IR
    inet = InpaintSANet(c_dim=4)
    isd = S.fill_state_dict(inet.state_dict(), seed=3, conv_std=0.05)
    inet.load_state_dict(isd)
    inet = inet.to(cuda).eval()
    isd_gpu = {k: v.to(cuda) for k, v in isd.items()}
    img1 = S.synthetic_source(size, seed=5).to(cuda)
    ys, xs = torch.meshgrid(torch.linspace(-1, 1, size), torch.linspace(-1, 1, size), indexing="ij")
    msk = (((xs / 0.4) ** 2 + (ys / 0.8) ** 2) < 1).float()[None, None].to(cuda)
    _, x_ref, _ = IR.forward(img1, msk, isd_gpu)
    x_lwb = inet(img1, msk, only_x=True)
    out["inpaintor_lwb_vs_stock_fp32_max_abs"] = (x_lwb - x_ref).abs().max().item()
    out["inpaintor_stock_ms_fp32"] = timeit(lambda: IR.forward(img1, msk, isd_gpu), iters=5, warm=2)
    out["inpaintor_lwb_ms"] = timeit(lambda: inet(img1, msk, only_x=True), iters=10, warm=3)
    assert out["inpaintor_lwb_vs_stock_fp32_max_abs"] < 1e-3
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "stock_compare.json"), "w"), indent=1)
    assert out["lwb_generator_ms"] < out["stock_generator_ms_fp32"]
