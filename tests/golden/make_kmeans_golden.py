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


def equalize_inputs(mode):
    """nn_index [40000] int64: cluster 2 has 15000 members, cluster 5 has 11000 (both above max_cnt_th = 10000),
    cluster 6 is empty, the rest share the remainder; leaf mode spreads over 8*3+1 ids incl. the sentinel 24."""
    rs = np.random.RandomState(3 if mode == "root" else 4)
    k = 8 if mode == "root" else 25
    big, second, empty = (2, 5, 6) if mode == "root" else (7, 24, 11)
    others = [i for i in range(k) if i not in (big, second, empty)]
    ids = np.concatenate([np.full(15000, big), np.full(11000, second), rs.choice(others, size=14000)])
    rs.shuffle(ids)
    return ids.astype(np.int64)


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
    # equalize_cluster_size (:89-144) on crafted assignments: two clusters above the 10000-member threshold (their
    # overflow goes to excl_cluster_ids), an empty cluster, and the leaf mode's k1*k2+1 rows with the sentinel id
    for mode in ("root", "leaf"):
        q = ref.Quantize_kMeans(num_clusters=8, num_leaf_clusters=3, num_iters=1, dim=9)
        q.nn_index = torch.from_numpy(equalize_inputs(mode))
        q.equalize_cluster_size(mode=mode)
        out[f"equalize_{mode}/cluster_ids"] = q.cluster_ids.numpy().astype(np.int32)
        out[f"equalize_{mode}/cluster_len"] = q.cluster_len.numpy().astype(np.int64)
        out[f"equalize_{mode}/max_cnt"] = np.int64(int(q.max_cnt))
        out[f"equalize_{mode}/excl_clusters"] = np.array([int(e) for e in q.excl_clusters], np.int64)
        for i, e in enumerate(q.excl_cluster_ids):
            out[f"equalize_{mode}/excl_cluster_ids/{i}"] = e.numpy().astype(np.int32)
        assert q.n_excl_cls == 2 and (q.cls_ids if mode == "root" else q.leaf_cls_ids) is q.nn_index
    np.savez_compressed(os.path.join(HERE, "kmeans_golden.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, v.dtype)


if __name__ == "__main__":
    main()
