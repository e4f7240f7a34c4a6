"""Generates tests/golden/teapot.npz -- run ONLY in the build container (needs /root/reference).

Reproduces the inputs/expected outputs of the reference's own rasterizer known-answer tests
  thirdparty/neural_renderer/tests/test_rasterize_silhouettes.py:16-35  (silhouette == Blender render, exact)
  thirdparty/neural_renderer/tests/test_rasterize_depth.py:37-54        (normalised depth, atol 1e-2)
by importing the reference's own look_at.py / perspective.py / vertices_to_faces.py (pure torch)
and parsing teapot.obj the way load_obj.py:100-147 does (that function itself calls .cuda()).
Stored: the transformed face tensor that reaches rasterize_cuda.forward_face_index_map, and the
two golden images.
"""
import importlib.util
import os
import sys

import cv2
import numpy as np
import torch

REF = "/root/reference/thirdparty/neural_renderer"
HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, "neural_renderer", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_obj(path):  # load_obj.py:100-147 with normalization=True, on CPU
    vertices, faces = [], []
    lines = open(path).readlines()
    for line in lines:
        sp = line.split()
        if len(sp) and sp[0] == "v":
            vertices.append([float(v) for v in sp[1:4]])
    for line in lines:
        sp = line.split()
        if len(sp) and sp[0] == "f":
            vs = sp[1:]
            v0 = int(vs[0].split("/")[0])
            for i in range(len(vs) - 2):
                faces.append((v0, int(vs[i + 1].split("/")[0]), int(vs[i + 2].split("/")[0])))
    vertices = torch.from_numpy(np.vstack(vertices).astype(np.float32))
    faces = torch.from_numpy(np.vstack(faces).astype(np.int32)) - 1
    vertices -= vertices.min(0)[0][None, :]
    vertices /= torch.abs(vertices).max()
    vertices *= 2
    vertices -= vertices.max(0)[0][None, :] / 2
    return vertices, faces


def main():
    import math
    look_at = _load("look_at").look_at
    perspective = _load("perspective").perspective
    vertices_to_faces = _load("vertices_to_faces").vertices_to_faces
    v, f = load_obj(os.path.join(REF, "tests/data/teapot.obj"))
    assert v.shape[0] == 1292 and f.shape[0] == 2464          # tests/test_load_obj.py:37-41
    v, f = v[None], f[None]
    f = torch.cat((f, f[:, :, [2, 1, 0]]), dim=1)              # renderer.py:77-78 fill_back=True
    eye = [0, 0, -(1. / math.tan(math.radians(30)) + 1)]        # renderer.py:42
    v = look_at(v, eye)
    v = perspective(v, angle=30)
    faces = vertices_to_faces(v, f)[0].numpy().astype(np.float32)
    sil = cv2.imread(os.path.join(REF, "tests/data/teapot_blender.png"), cv2.IMREAD_UNCHANGED)
    sil = (sil[..., :3].min(-1) != 255) if sil.ndim == 3 else (sil != 255)
    dep = cv2.imread(os.path.join(REF, "tests/data/test_depth.png"), cv2.IMREAD_UNCHANGED)
    if dep.ndim == 3:
        dep = dep[..., 0]
    np.savez_compressed(os.path.join(HERE, "teapot.npz"), faces=faces,
                        silhouette=np.packbits(sil.astype(np.uint8)), depth_u8=dep.astype(np.uint8))
    print("faces", faces.shape, "covered", int(sil.sum()))


if __name__ == "__main__":
    sys.exit(main())
