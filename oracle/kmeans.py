"""ctypes front-end of oracle/kmeans_oracle.c + a numpy mirror of the control flow of
``Quantize_kMeans.cluster_assign`` (reference scene/kmeans_quantize.py:146-241).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pinned against the reference file itself:
tests/golden/make_kmeans_golden.py runs /root/reference/scene/kmeans_quantize.py on the CPU and
commits centres + ids; tests/test_oracle_cpu.py checks this mirror against those fixtures.
"""
import ctypes as C

import numpy as np

from . import lib

CHUNK = 10000


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def assign(a, b, scale_b, centers, select_ids=None, selected=-1, id_offset=0, ids_out=None):
    L = lib()
    a = np.ascontiguousarray(a, np.float32)
    b = None if b is None else np.ascontiguousarray(b, np.float32)
    centers = np.ascontiguousarray(centers, np.float32)
    N, Da = a.shape
    Db = 0 if b is None else b.shape[1]
    if ids_out is None:
        ids_out = np.zeros(N, np.int64)
    sel = None if select_ids is None else np.ascontiguousarray(select_ids, np.int64)
    rc = L.ogs_oracle_kmeans_assign(C.c_int64(N), _p(a), C.c_int(Da), _p(b), C.c_int(Db), C.c_float(scale_b),
                                    _p(centers), C.c_int(centers.shape[0]), _p(sel), C.c_int64(selected),
                                    C.c_int64(id_offset), _p(ids_out))
    assert rc == 0
    return ids_out


def accumulate(a, b, scale_b, k, ids, select_ids=None, selected=-1, id_offset=0, group=256):
    L = lib()
    a = np.ascontiguousarray(a, np.float32)
    b = None if b is None else np.ascontiguousarray(b, np.float32)
    N, Da = a.shape
    Db = 0 if b is None else b.shape[1]
    sums = np.zeros((k, Da + Db), np.float32)
    counts = np.zeros(k, np.float32)
    sel = None if select_ids is None else np.ascontiguousarray(select_ids, np.int64)
    rc = L.ogs_oracle_kmeans_accumulate(C.c_int64(N), _p(a), C.c_int(Da), _p(b), C.c_int(Db), C.c_float(scale_b),
                                        C.c_int(k), _p(np.ascontiguousarray(ids, np.int64)), _p(sel),
                                        C.c_int64(selected), C.c_int64(id_offset), C.c_int(group), _p(sums), _p(counts))
    assert rc == 0
    return sums, counts


def cluster_assign_root(a, b, scale_b, centers0, num_iters):
    """Mirror of the root branch of cluster_assign: returns (centres [k,D], ids [N])."""
    N = a.shape[0]
    k = centers0.shape[0]
    centers = np.asarray(centers0, np.float32).copy()
    counts_state = np.full(k, 1e-6, np.float32)
    n_eps = N // CHUNK + 1
    for _ in range(num_iters):
        ids = assign(a, b, scale_b, centers)
        sums, cnt = accumulate(a, b, scale_b, k, ids)
        counts_state = (counts_state + (cnt + np.float32(n_eps * 1e-6))).astype(np.float32)
        centers = (sums / counts_state[:, None]).astype(np.float32)
        counts_state[counts_state > 0.1] = 0.0
    ids = assign(a, b, scale_b, centers)
    return centers, ids


def cluster_assign_leaf(feat, cls_ids, leaf_centers0, leaf_cls_ids0, selected, n_sub, k2, num_iters):
    """Mirror of the leaf branch: returns (leaf_centres, leaf_cls_ids)."""
    leaf_centers = np.asarray(leaf_centers0, np.float32).copy()
    ids = np.asarray(leaf_cls_ids0, np.int64).copy()
    start = selected * k2
    D = feat.shape[1]
    counts_state = np.full(k2, 1e-6, np.float32)
    for _ in range(num_iters):
        assign(feat, None, 1.0, leaf_centers[start:start + n_sub], cls_ids, selected, start, ids)
        s, c = accumulate(feat, None, 1.0, n_sub, ids, cls_ids, selected, start)
        sums = np.zeros((k2, D), np.float32)
        cnt = np.zeros(k2, np.float32)
        sums[:n_sub] = s
        cnt[:n_sub] = c
        counts_state = (counts_state + (cnt + np.float32(1e-6))).astype(np.float32)
        leaf_centers[start:start + k2] = (sums / counts_state[:, None]).astype(np.float32)
        counts_state[counts_state > 0.1] = 0.0
    assign(feat, None, 1.0, leaf_centers[start:start + n_sub], cls_ids, selected, start, ids)
    return leaf_centers, ids


def assign_segmented(a, coarse_ids, seg_centers, seg_k, k2, ids_out=None):
    """All coarse clusters' leaf assigns (scene/kmeans_quantize.py:196-206 run for every idx_c): k1 calls of the
    per-cluster oracle, each restricted to its own points and its own centre block."""
    a = np.ascontiguousarray(a, np.float32)
    k1 = len(seg_k)
    if ids_out is None:
        ids_out = np.full(a.shape[0], k1 * k2, np.int64)
    for c in range(k1):
        n = int(min(max(int(seg_k[c]), 0), k2))
        if n > 0:
            assign(a, None, 1.0, seg_centers[c * k2:c * k2 + n], coarse_ids, c, c * k2, ids_out)
    return ids_out


def accumulate_fixed(a, ids, rows, fix_bits, valid):
    """Exact fixed-point centroid sums and counts: int64 [rows, D+1] with acc[r, d] = sum rint(x_d * 2^fix_bits),
    acc[r, D] = member count, over the points flagged `valid`."""
    a = np.ascontiguousarray(a, np.float32)
    D = a.shape[1]
    q = np.rint(a.astype(np.float64) * float(2 ** fix_bits)).astype(np.int64)
    acc = np.zeros((rows, D + 1), np.int64)
    sel = np.flatnonzero(valid)
    for d in range(D):
        np.add.at(acc[:, d], ids[sel], q[sel, d])
    np.add.at(acc[:, D], ids[sel], 1)
    return acc
