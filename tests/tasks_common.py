"""Shared set-up of the Viewer / Swapper parity tests: the inputs tests/golden/make_tasks_golden.py used."""
import os

import numpy as np
import torch

from impersonator_b200 import synthetic as S

GOLD = os.path.join(os.path.dirname(__file__), "golden", "tasks.npz")
SIZE = 128


def sl(t):
    return t[:, :, 1::4, 2::4].contiguous().cpu().numpy()


def write_inputs(folder):
    a, b = os.path.join(str(folder), "a.png"), os.path.join(str(folder), "b.png")
    S.save_png(S.synthetic_source(160, seed=71)[0], a)
    S.save_png(S.synthetic_source(SIZE, seed=72)[0], b)
    return a, b


def read_like_reference(path, size=SIZE):
    """cv_utils.read_cv2_img + transform_img(transpose=True) * 2 - 1 (utils/cv_utils.py:10-47, models/viewer.py:85-89)."""
    import cv2
    img = cv2.cvtColor(cv2.imread(path, -1), cv2.COLOR_BGR2RGB)
    x = cv2.resize(img, (size, size)).astype(np.float32) / 255.0
    return torch.from_numpy(x.transpose((2, 0, 1)) * 2 - 1.0)[None].float()


def part_table(nf):
    info = S.synthetic_part_info()
    names = sorted(info.keys())
    fn = torch.zeros(nf + 1, len(names) + 1)
    for i, name in enumerate(names):
        fn[info[name]["face"], i] = 1.0
    fn[-1, -1] = 1.0
    return info, fn, [info[name]["face"] for name in names]


class Opt(object):
    image_size, batch_size, bg_model, repeat_num, cond_nc = SIZE, 4, "ORIGINAL", 6, 3
    bg_ks, ft_ks, front_warp, only_vis, bg_replace = 13, 3, False, False, False
