"""Generates tests/golden/kmeans_golden.npz by running the REFERENCE file itself
(/root/reference/scene/kmeans_quantize.py, loaded by path, torch CPU, `.cuda()` patched to a no-op).
Run in the build container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_kmeans_golden.py

Inputs are regenerated from numpy's MT19937 stream (stable across versions) by `inputs()` below,
so only the reference's OUTPUTS are stored.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/scene/kmeans_quantize.py"

CASES = {
    # name: (N, k1, k2, iters, pos_weight)
    "root_25k": (25_000, 64, 10, 5, 1.0),
    "root_20k_exact_chunks": (20_000, 32, 5, 3, 0.5),
}


def inputs(name):
    N, k1, k2, iters, pw = CASES[name]
    rs = np.random.RandomState({"root_25k": 11, "root_20k_exact_chunks": 12}[name])
    # clustered features so that k-means has structure: 40 blobs in the 6-D unit cube + xyz in a room
    blobs = rs.rand(40, 6).astype(np.float32)
    which = rs.randint(0, 40, size=N)
    ins_feat = (blobs[which] + 0.05 * rs.randn(N, 6)).astype(np.float32)
    xyz = ((rs.rand(N, 3) - 0.5) * np.array([8.0, 6.0, 3.0])).astype(np.float32)
    return ins_feat, xyz


def load_reference():
    torch.Tensor.cuda = lambda self, *a, **k: self
    for m in ("tqdm",):
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = types.SimpleNamespace(tqdm=lambda x, **k: x)
    spec = importlib.util.spec_from_file_location("ref_kmeans_quantize", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class G:
    pass


def main():
    torch.set_num_threads(1)   # fixed BLAS reduction order for the fixture
    ref = load_reference()
    out = {}
    for name, (N, k1, k2, iters, pw) in CASES.items():
        ins_feat, xyz = inputs(name)
        g = G()
        g._ins_feat = torch.from_numpy(ins_feat.copy()).requires_grad_(True)
        g._xyz = torch.from_numpy(xyz.copy())
        q = ref.Quantize_kMeans(num_clusters=k1, num_leaf_clusters=k2, num_iters=iters, dim=9)
        feat0 = np.concatenate([ins_feat, xyz * np.float32(pw)], 1)
        q.centers = torch.from_numpy(feat0[:k1].copy())          # injected: bypasses torch.randperm
        q.forward(g, 1, assign=True, mode="root", pos_weight=pw)
        out[f"{name}/centers"] = q.centers.numpy().copy()
        out[f"{name}/cls_ids"] = q.cls_ids.numpy().astype(np.int16)
        out[f"{name}/ins_feat_q"] = g._ins_feat_q.detach().numpy()[:512].copy()
        if name == "root_25k":
            # fine level for two coarse clusters, one of them with fewer sub-clusters than k2
            q.leaf_centers = torch.from_numpy(ins_feat[:k1 * k2 + 1].copy())
            q.leaf_cls_ids = torch.ones(N).to(torch.int64) * k1 * k2
            sub = torch.full((k1,), k2, dtype=torch.int64)
            sub[7] = 4
            q.iLeafSubNum = sub
            for sel in (3, 7):
                q.forward(g, 1, assign=True, mode="leaf", selected_leaf=sel)
            out[f"{name}/leaf_centers"] = q.leaf_centers.numpy().copy()
            out[f"{name}/leaf_cls_ids"] = q.leaf_cls_ids.numpy().astype(np.int16)
            out[f"{name}/leaf_ins_feat_q"] = g._ins_feat_q.detach().numpy()[:512].copy()
    np.savez_compressed(os.path.join(HERE, "kmeans_golden.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, v.dtype)


if __name__ == "__main__":
    main()
