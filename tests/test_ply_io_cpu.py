"""PLY checkpoint layout against the vertex table the REFERENCE's own save_ply / load_ply produced
(tests/golden/make_ply_golden.py, scene/gaussian_model.py:249-351)."""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from opengaussian_b200 import ply_io  # noqa: E402

from make_ply_golden import inputs  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "ply_golden.npz"))
KEYS = ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation", "_ins_feat")


def _model():
    return types.SimpleNamespace(**inputs(), max_sh_degree=3)


def test_attribute_order_matches_reference():
    assert ply_io.attribute_names() + ["red", "green", "blue"] == list(GOLD["plain/names"])


def test_vertex_table_is_byte_identical(tmp_path):
    for tag, save_q in (("plain", ()), ("quantised", ("ins_feat",))):
        pc = _model()
        path = str(tmp_path / tag / "point_cloud.ply")
        ply_io.save_ply(path, pc, save_q)
        el = ply_io.read_ply(path)
        assert list(el.dtype.names) == list(GOLD[f"{tag}/names"])
        assert [el.dtype[n].str for n in el.dtype.names] == list(GOLD[f"{tag}/formats"])
        assert el.tobytes() == GOLD[f"{tag}/bytes"].tobytes()
        raw = open(path, "rb").read()
        assert raw.startswith(b"ply\nformat binary_little_endian 1.0\nelement vertex 257\nproperty float x\n")
        assert raw.endswith(el.tobytes())


def test_load_matches_reference_loader(tmp_path):
    for tag, save_q in (("plain", ()), ("quantised", ("ins_feat",))):
        path = str(tmp_path / f"{tag}.ply")
        ply_io.save_ply(path, _model(), save_q)
        got = ply_io.load_ply(path, 3, device="cpu")
        for k in KEYS:
            ref = GOLD[f"{tag}/loaded{k}"]
            assert tuple(got[k].shape) == ref.shape, k
            assert got[k].is_contiguous() and got[k].dtype == torch.float32
            assert np.array_equal(got[k].numpy(), ref), k


def test_round_trip_and_preview_colour(tmp_path):
    pc = _model()
    pc._ins_feat[0] = torch.tensor([3.0, -3.0, 0.0, 0, 0, 0])      # clipped to 255 / 0, mid -> 127
    pc._opacity[0] = 5.0
    pc._opacity[1] = -5.0                                          # sigmoid < 0.1 -> grey
    path = str(tmp_path / "a.ply")
    ply_io.save_ply(path, pc)
    el = ply_io.read_ply(path)
    assert (el["red"][0], el["green"][0], el["blue"][0]) == (255, 0, 127)
    assert (el["red"][1], el["green"][1], el["blue"][1]) == (128, 128, 128)
    assert not el["nx"].any() and not el["ny"].any() and not el["nz"].any()
    got = ply_io.load_ply(path, 3, device="cpu")
    for k in KEYS:
        assert torch.equal(got[k], getattr(pc, k)), k


def test_reader_accepts_ascii_and_other_orders(tmp_path):
    path = str(tmp_path / "b.ply")
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment hand written\nelement vertex 2\nproperty double z\nproperty float x\n"
                "property uchar red\nend_header\n1.5 2 7\n-1 0.25 255\n")
    el = ply_io.read_ply(path)
    assert el.dtype.names == ("z", "x", "red")
    assert el["z"].tolist() == [1.5, -1.0] and el["x"].tolist() == [2.0, 0.25] and el["red"].tolist() == [7, 255]


def test_empty_cloud(tmp_path):
    pc = types.SimpleNamespace(**{k: v[:0] for k, v in inputs().items()})
    path = str(tmp_path / "e.ply")
    ply_io.save_ply(path, pc)
    got = ply_io.load_ply(path, 3, device="cpu")
    assert got["_xyz"].shape == (0, 3) and got["_features_rest"].shape == (0, 15, 3)
