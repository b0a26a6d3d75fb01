"""Generates tests/golden/ply_golden.npz by running the REFERENCE's own GaussianModel.construct_list_of_attributes /
save_ply / load_ply (scene/gaussian_model.py:249-298, :305-351; the three method definitions and `sigmoid` are
taken from the source with `ast` -- the module itself needs plyfile, simple_knn and a CUDA device).

The third-party `plyfile` package is not installed here.  A stand-in captures the structured vertex array that
save_ply hands to `PlyElement.describe` (that array IS the file's payload) and serves load_ply from such an array;
`torch.tensor(..., device="cuda")` is redirected to the CPU.  Stored: the vertex table's bytes / field names and
the tensors the reference's load_ply produced from it.      Run:  python tests/golden/make_ply_golden.py
"""
import ast
import os
import types

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/scene/gaussian_model.py"


def inputs(P=257, seed=5):
    rs = np.random.RandomState(seed)
    f = lambda *s: torch.from_numpy(rs.randn(*s).astype(np.float32))  # noqa: E731
    return dict(_xyz=f(P, 3), _features_dc=f(P, 1, 3), _features_rest=f(P, 15, 3), _opacity=f(P, 1) * 3,
                _scaling=f(P, 3), _rotation=f(P, 4), _ins_feat=f(P, 6) * 0.8, _ins_feat_q=f(P, 6) * 0.8)


class _Captured:
    array = None


class PlyElement:
    @staticmethod
    def describe(elements, name):
        assert name == "vertex"
        _Captured.array = elements.copy()
        return elements


class _Prop:
    def __init__(self, name):
        self.name = name


class _Element:
    def __init__(self, arr):
        self.arr = arr
        self.properties = [_Prop(n) for n in arr.dtype.names]

    def __getitem__(self, name):
        return self.arr[name]


class PlyData:
    def __init__(self, els=None):
        self.elements = [] if els is None else [_Element(e) for e in els]

    def write(self, path):
        pass

    @staticmethod
    def read(path):
        return PlyData([_Captured.array])


def load_reference():
    tree = ast.parse(open(REF).read())
    real_tensor = torch.tensor
    ns = {"torch": torch, "np": np, "nn": nn, "os": os, "mkdir_p": lambda p: None, "PlyElement": PlyElement,
          "PlyData": PlyData}
    fns = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "sigmoid":
            exec(compile(ast.Module(body=[node], type_ignores=[]), REF, "exec"), ns)
        if isinstance(node, ast.ClassDef) and node.name == "GaussianModel":
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name in ("construct_list_of_attributes", "save_ply", "load_ply"):
                    exec(compile(ast.Module(body=[item], type_ignores=[]), REF, "exec"), ns)
                    fns[item.name] = ns[item.name]
    ns["torch"] = types.SimpleNamespace(**{k: getattr(torch, k) for k in dir(torch) if not k.startswith("__")})
    ns["torch"].tensor = lambda *a, **k: real_tensor(*a, **{kk: vv for kk, vv in k.items() if kk != "device"})
    return fns


def main():
    fns = load_reference()
    out = {}
    for tag, save_q in (("plain", []), ("quantised", ["ins_feat"])):
        inp = inputs()
        me = types.SimpleNamespace(**inp, max_sh_degree=3)
        me.construct_list_of_attributes = lambda me=me: fns["construct_list_of_attributes"](me)
        fns["save_ply"](me, "/tmp/unused/point_cloud.ply", save_q=save_q)
        arr = _Captured.array
        out[f"{tag}/names"] = np.array(arr.dtype.names)
        out[f"{tag}/formats"] = np.array([arr.dtype[n].str for n in arr.dtype.names])
        out[f"{tag}/bytes"] = np.frombuffer(arr.tobytes(), dtype=np.uint8)
        loaded = types.SimpleNamespace(max_sh_degree=3)
        fns["load_ply"](loaded, "/tmp/unused/point_cloud.ply")
        for k in ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation", "_ins_feat"):
            out[f"{tag}/loaded{k}"] = getattr(loaded, k).detach().numpy()
    path = os.path.join(HERE, "ply_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, out["plain/names"][:14], out["plain/bytes"].shape)


if __name__ == "__main__":
    main()
