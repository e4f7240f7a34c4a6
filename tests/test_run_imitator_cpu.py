"""tests/run_imitator_body.py must be the reference's run_imitator.py main block, character for character (checked where the
reference tree exists; the GPU box runs the committed copy)."""
import os

import pytest

from run_imitator_body import BODY

REF = "/root/reference/run_imitator.py"


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present (GPU box)")
def test_body_is_the_reference_main_block():
    src = open(REF).read()
    main = src[src.index('if __name__ == "__main__":'):]
    main = main[main.index("\n") + 1:]
    assert main.rstrip("\n") == BODY.rstrip("\n")


def test_renderer_tables_from_asset_files(tmp_path, monkeypatch):
    """SMPLRenderer(image_size, tex_size, has_front, fill_back=False) as models/imitator.py:39-41 calls it: faces + lookup
    tables come from the files under assets/pretrains (synthetic files in the real formats here)."""
    import numpy as np
    from impersonator_b200 import mesh, synthetic as S
    from impersonator_b200.nmr import SMPLRenderer
    S.write_synthetic_assets(str(tmp_path), n_targets=1)
    monkeypatch.chdir(tmp_path)
    r = SMPLRenderer(image_size=256, tex_size=3, has_front=True, fill_back=False)
    assert tuple(r.faces.shape) == (13776, 3) and tuple(r.map_fn.shape) == (13777, 3)
    assert r.map_fn[-1].tolist() == [0.0, 0.0, 1.0] and float(r.map_fn[:-1, 2].abs().max()) == 0.0
    assert float(r.front_map_fn.sum()) == 500.0 and float(r.back_map_fn.sum()) == 700.0      # head minus front
    assert mesh.get_map_fn_dim('uv_seg') == 3
    if os.path.exists("/root/reference/utils/mesh.py"):
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_mesh", "/root/reference/utils/mesh.py")
        R = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(R)
        for name in ("uv_seg", "front", "back", "head", "uv", "seg"):
            for fb in (False, True):
                mine = mesh.create_mapping(name, "assets/pretrains/mapper.txt", contain_bg=True, fill_back=fb)
                ref = R.create_mapping(name, "assets/pretrains/mapper.txt", contain_bg=True, fill_back=fb)
                assert mine.dtype == ref.dtype and np.array_equal(mine, ref), (name, fb)
        for fb in (False, True):                                  # the Swapper's tables (models/swapper.py:34-37)
            if fb:                                                # upstream's 'par' table does not support fill_back either
                for fn in (mesh.create_mapping, R.create_mapping):
                    with pytest.raises(AssertionError):
                        fn('par', "assets/pretrains/mapper.txt", contain_bg=True, fill_back=True)
            else:
                mine = mesh.create_mapping('par', "assets/pretrains/mapper.txt", contain_bg=True, fill_back=False)
                ref = R.create_mapping('par', "assets/pretrains/mapper.txt", contain_bg=True, fill_back=False)
                assert mine.shape == ref.shape == (13776 + 1, 11) and np.array_equal(mine, ref)
            mine_ids = mesh.get_part_face_ids('par', "assets/pretrains/mapper.txt", fill_back=fb)
            ref_ids = R.get_part_face_ids('par', "assets/pretrains/mapper.txt", fill_back=fb)
            assert list(mine_ids.keys()) == list(ref_ids.keys())
            assert all(list(mine_ids[k]) == list(ref_ids[k]) for k in ref_ids), fb
            for kind in ('head_front', 'head_back'):
                assert sorted(mesh.get_part_face_ids(kind, "assets/pretrains/mapper.txt", fill_back=fb)) == \
                    sorted(R.get_part_face_ids(kind, "assets/pretrains/mapper.txt", fill_back=fb)), (kind, fb)


def test_networks_factory_names():
    from impersonator_b200.networks import NetworksFactory, HumanModelRecovery       # noqa: F401
    import pytest as _pt
    with _pt.raises(ValueError):
        NetworksFactory.get_by_name('nope')
    g = NetworksFactory.get_by_name('impersonator', bg_dim=4, src_dim=6, tsf_dim=6, repeat_num=6)
    assert g.n_down == 3 and g.repeat_num == 6
