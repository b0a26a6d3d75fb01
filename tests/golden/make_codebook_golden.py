"""Generates tests/golden/codebook_golden.npz by running the REFERENCE's own save / load code:
train.py::dec2binary + save_kmeans (:52-100, taken from the source with `ast`: train.py itself cannot be
imported here) and utils/opengs_utlis.py::load_code_book (:68-88, module loaded by path).

Both need the third-party `bitarray` package, which is not installed in this container.  A minimal stand-in
implementing the calls they make (constructor from a list of bools, extend, len, tofile, fromfile, slicing,
tolist) with bitarray's documented default bit order (big-endian within each byte, zero padding to a whole
byte) is injected; everything else is the reference's code.  Stored: the three files the reference wrote
(bytes) and the indices its loader returned.   Run:  python tests/golden/make_codebook_golden.py
"""
import ast
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


class bitarray:                                       # noqa: N801  (the third-party class name)
    def __init__(self, init=None):
        self.bits = [] if init is None else [bool(b) for b in init]

    def extend(self, other):
        self.bits.extend(other.bits)

    def __len__(self):
        return len(self.bits)

    def tofile(self, f):
        f.write(np.packbits(np.asarray(self.bits, dtype=np.uint8), bitorder="big").tobytes())

    def fromfile(self, f):
        self.bits.extend(bool(b) for b in np.unpackbits(np.frombuffer(f.read(), dtype=np.uint8), bitorder="big"))

    def __getitem__(self, s):
        out = bitarray()
        out.bits = self.bits[s]
        return out

    def tolist(self):
        return list(self.bits)


CASES = {"root_1000": (1000, 64, "root"), "leaf_777": (777, 641, "leaf")}   # 777 * 10 bits: not a whole byte


def inputs(name):
    n, k, mode = CASES[name]
    rs = np.random.RandomState(len(name))
    ids = torch.from_numpy(rs.randint(0, k, size=n).astype(np.int64))
    centers = torch.from_numpy(rs.rand(k, 6).astype(np.float32))
    return ids, centers, mode


def load_reference():
    sys.modules["bitarray"] = types.SimpleNamespace(bitarray=bitarray)
    spec = importlib.util.spec_from_file_location("ref_opengs_utlis", os.path.join(REF, "utils/opengs_utlis.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    tree = ast.parse(open(os.path.join(REF, "train.py")).read())
    ns = {"torch": torch, "np": np, "os": os, "bitarray": bitarray, "mkdir_p": lambda p: os.makedirs(p, exist_ok=True)}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("dec2binary", "save_kmeans"):
            exec(compile(ast.Module(body=[node], type_ignores=[]), "train.py", "exec"), ns)
    return mod, ns["save_kmeans"]


def main():
    ref, save_kmeans = load_reference()
    out = {}
    for name in CASES:
        ids, centers, mode = inputs(name)
        km = types.SimpleNamespace(cls_ids=ids, leaf_cls_ids=ids, centers=centers, leaf_centers=centers)
        with tempfile.TemporaryDirectory() as d:
            save_kmeans([km], ["ins_feat"], d, mode=mode)
            sub = os.path.join(d, f"{mode}_code_book")
            out[f"{name}/inds"] = np.fromfile(os.path.join(sub, "kmeans_inds.bin"), dtype=np.uint8)
            args = np.load(os.path.join(sub, "kmeans_args.npy"), allow_pickle=True).item()
            out[f"{name}/n_bits"] = np.int64(args["n_bits"])
            out[f"{name}/total_len"] = np.int64(args["total_len"])
            codebook, loaded = ref.load_code_book(sub)
            out[f"{name}/loaded"] = np.asarray(loaded, dtype=np.int64)
            out[f"{name}/centers"] = codebook["ins_feat"].numpy()
    path = os.path.join(HERE, "codebook_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
