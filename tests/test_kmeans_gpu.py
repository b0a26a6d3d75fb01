"""GPU parity of the k-means codebook kernels (through the Quantize_kMeans drop-in / C ABI).

* one assign pass: ids BIT-EXACT vs the C oracle (same fmaf chain, lowest-index ties);
  fused centroid sums/counts: counts exact, sums within fp32 reordering (1e-5 relative);
* full cluster_assign (root + leaf) vs the golden vectors produced by the reference file itself:
  ids equal away from near-ties (<= 2e-3 mismatches), centres within 2e-3;
* properties at BASELINE size (5 M points): every id is the argmin under a recomputed fp64
  distance up to fp32 slack; counts sum to N; idempotence of the reassign.
"""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import kmeans as okm

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "kmeans_golden.npz")


def _golden_inputs(name):
    spec = importlib.util.spec_from_file_location("mkg", os.path.join(os.path.dirname(GOLD), "make_kmeans_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.inputs(name), m.CASES[name]


class _G:
    pass


@pytest.mark.parametrize("N,Da,Db,k", [(1, 6, 3, 64), (31, 6, 0, 10), (10_000, 6, 3, 64), (123_457, 6, 3, 64),
                                       (50_000, 6, 0, 7), (4096, 3, 0, 1), (70_000, 9, 0, 200)])
def test_assign_bit_exact_and_fused_sums(N, Da, Db, k):
    from opengaussian_b200.kmeans_quantize import kmeans_assign
    rs = np.random.RandomState(N % 97)
    a = rs.rand(N, Da).astype(np.float32)
    b = (rs.rand(N, Db).astype(np.float32) * 4 - 2) if Db else None
    centers = rs.rand(k, Da + Db).astype(np.float32)
    if k > 2:
        centers[2] = centers[1]                       # duplicate centre: ties must go to the lower index
    scale_b = 0.37
    want = okm.assign(a, b, scale_b, centers)
    sums = torch.zeros(k, Da + Db, device="cuda")
    cnt = torch.zeros(k, device="cuda")
    ids = kmeans_assign(torch.from_numpy(a).cuda(), None if b is None else torch.from_numpy(b).cuda(), scale_b,
                        torch.from_numpy(centers).cuda(), sums=sums, counts=cnt)
    assert ids.dtype == torch.int64
    assert np.array_equal(ids.cpu().numpy(), want)
    wsum, wcnt = okm.accumulate(a, b, scale_b, k, want)
    assert np.array_equal(cnt.cpu().numpy(), wcnt)
    assert np.allclose(sums.cpu().numpy(), wsum, rtol=1e-5, atol=1e-4 * max(1.0, np.abs(wsum).max() * 1e-2))


def test_assign_select_and_offset():
    from opengaussian_b200.kmeans_quantize import kmeans_assign
    rs = np.random.RandomState(5)
    N = 30_000
    a = rs.rand(N, 6).astype(np.float32)
    sel = rs.randint(0, 8, size=N).astype(np.int64)
    centers = rs.rand(5, 6).astype(np.float32)
    out0 = np.full(N, 640, np.int64)
    want = okm.assign(a, None, 1.0, centers, sel, 3, 30, out0.copy())
    ids = torch.from_numpy(out0.copy()).cuda()
    sums = torch.zeros(5, 6, device="cuda")
    cnt = torch.zeros(5, device="cuda")
    kmeans_assign(torch.from_numpy(a).cuda(), None, 1.0, torch.from_numpy(centers).cuda(),
                  torch.from_numpy(sel).cuda(), 3, 30, ids, sums, cnt)
    assert np.array_equal(ids.cpu().numpy(), want)
    assert int(cnt.sum()) == int((sel == 3).sum())


def test_quantize_kmeans_vs_reference_golden():
    from opengaussian_b200.kmeans_quantize import Quantize_kMeans
    gold = np.load(GOLD)
    for name in ("root_25k", "root_20k_exact_chunks"):
        (ins_feat, xyz), (N, k1, k2, iters, pw) = _golden_inputs(name)
        g = _G()
        g._ins_feat = torch.from_numpy(ins_feat).cuda().requires_grad_(True)
        g._xyz = torch.from_numpy(xyz).cuda()
        q = Quantize_kMeans(num_clusters=k1, num_leaf_clusters=k2, num_iters=iters, dim=9)
        q.centers = torch.from_numpy(np.concatenate([ins_feat, xyz * np.float32(pw)], 1)[:k1]).cuda()
        q.forward(g, 1, assign=True, mode="root", pos_weight=pw)
        ref_ids = gold[f"{name}/cls_ids"].astype(np.int64)
        assert q.cls_ids.dtype == torch.int64 and q.centers.dtype == torch.float32
        assert (q.cls_ids.cpu().numpy() != ref_ids).mean() <= 2e-3
        assert np.allclose(q.centers.cpu().numpy(), gold[f"{name}/centers"], rtol=2e-3, atol=2e-3)
        fq = g._ins_feat_q.detach().cpu().numpy()[:512]
        assert (np.abs(fq - gold[f"{name}/ins_feat_q"]).max(1) > 1e-3).mean() <= 0.01
        # straight-through: gradient of _ins_feat_q flows unchanged into _ins_feat
        g._ins_feat_q.sum().backward()
        assert torch.equal(g._ins_feat.grad, torch.ones_like(g._ins_feat))
        # the CUDA path and the CPU oracle mirror agree closely
        oc, oi = okm.cluster_assign_root(ins_feat, xyz, pw, np.concatenate([ins_feat, xyz * np.float32(pw)], 1)[:k1], iters)
        assert (q.cls_ids.cpu().numpy() != oi).mean() <= 1e-4
        assert np.allclose(q.centers.cpu().numpy(), oc, rtol=1e-4, atol=1e-5)
        if name == "root_25k":
            q.cls_ids = torch.from_numpy(ref_ids).cuda()
            q.leaf_centers = torch.from_numpy(ins_feat[:k1 * k2 + 1].copy()).cuda()
            q.leaf_cls_ids = torch.ones(N, device="cuda").to(torch.int64) * k1 * k2
            sub = torch.full((k1,), k2, dtype=torch.int64)
            sub[7] = 4
            q.iLeafSubNum = sub
            for sel in (3, 7):
                q.forward(g, 1, assign=True, mode="leaf", selected_leaf=sel)
            ref_leaf = gold[f"{name}/leaf_cls_ids"].astype(np.int64)
            got = q.leaf_cls_ids.cpu().numpy()
            assert (got != ref_leaf).mean() <= 2e-3
            assert np.array_equal(got == k1 * k2, ref_leaf == k1 * k2)
            lc = q.leaf_centers.cpu().numpy()
            assert np.allclose(lc, gold[f"{name}/leaf_centers"], rtol=2e-3, atol=2e-3)
            assert np.all(lc[7 * k2 + 4:8 * k2] == 0.0)
            # lazily built equalize_cluster_size products keep the reference's shapes
            assert q.cluster_ids.numel() == (k1 * k2 + 1) * int(q.max_cnt)
            assert q.cluster_len.shape == (k1 * k2 + 1, 1)


def test_non_assign_forward_is_noop_on_centres():
    from opengaussian_b200.kmeans_quantize import Quantize_kMeans
    (ins_feat, xyz), (N, k1, k2, iters, pw) = _golden_inputs("root_20k_exact_chunks")
    g = _G()
    g._ins_feat = torch.from_numpy(ins_feat).cuda().requires_grad_(True)
    g._xyz = torch.from_numpy(xyz).cuda()
    q = Quantize_kMeans(num_clusters=k1, num_leaf_clusters=k2, num_iters=2, dim=9)
    q.forward(g, 1, assign=True, mode="root", pos_weight=pw)      # random init path (torch.randperm)
    c0, i0 = q.centers.clone(), q.nn_index.clone()
    q.forward(g, 2, assign=False, mode="root", pos_weight=pw)
    assert torch.equal(q.centers, c0) and torch.equal(q.nn_index, i0)
    assert torch.allclose(g._ins_feat_q, c0[i0][:, :6])


def test_full_size_properties():
    """BASELINE config 5 size (5 M points, coarse k=64, D=9): size-independent properties."""
    from opengaussian_b200.kmeans_quantize import kmeans_assign
    N, k = 5_000_000, 64
    gen = torch.Generator(device="cuda").manual_seed(0)
    a = torch.rand(N, 6, device="cuda", generator=gen)
    b = (torch.rand(N, 3, device="cuda", generator=gen) - 0.5) * 8
    centers = torch.cat([a[:k], b[:k] * 0.5], 1).contiguous()
    sums = torch.zeros(k, 9, device="cuda")
    cnt = torch.zeros(k, device="cuda")
    ids = kmeans_assign(a, b, 0.5, centers, sums=sums, counts=cnt)
    assert int(cnt.sum()) == N and int(ids.min()) >= 0 and int(ids.max()) < k
    assert torch.equal(torch.bincount(ids, minlength=k).float(), cnt)
    x = torch.cat([a, b * 0.5], 1)
    idx = torch.randint(0, N, (200_000,), device="cuda", generator=gen)
    d = torch.cdist(x[idx].double(), centers.double())
    best = d.min(1).values
    mine = d.gather(1, ids[idx, None])[:, 0]
    assert float((mine - best).max()) <= 1e-5            # chosen centre is the (fp32-slack) nearest
    assert torch.allclose(sums.double(), torch.zeros(k, 9, device="cuda", dtype=torch.float64).index_add_(0, ids, x.double()),
                          rtol=1e-4, atol=1e-1)
    # idempotence: same centres -> same ids
    assert torch.equal(kmeans_assign(a, b, 0.5, centers), ids)


@pytest.mark.parametrize("N,D,k1,k2", [(200_003, 6, 8, 5), (1, 6, 3, 2), (50_000, 3, 64, 10), (4099, 5, 2, 7)])
def test_segmented_assign_bit_exact_and_exact_sums(N, D, k1, k2):
    """ogs_kmeans_assign_segmented: ids bit-exact against k1 per-cluster calls of the oracle (ragged seg_k incl. 0,
    coarse ids outside [0, k1) untouched); the fixed-point centroid sums are EXACT integers."""
    from opengaussian_b200.kmeans_quantize import kmeans_assign_segmented
    rs = np.random.RandomState(N % 89)
    a = (rs.rand(N, D).astype(np.float32) * 2 - 1)
    coarse = rs.randint(-1, k1 + 1, size=N).astype(np.int64)          # -1 and k1 are "not mine"
    seg_k = rs.randint(0, k2 + 1, size=k1).astype(np.int32)
    seg_k[0] = k2
    if k1 > 2:
        seg_k[1] = 0
    centers = rs.rand(k1 * k2 + 1, D).astype(np.float32) * 2 - 1
    if k2 > 2:
        centers[1] = centers[0]                                        # tie: lowest index wins
    want = okm.assign_segmented(a, coarse, centers, seg_k, k2)
    fix = 30
    acc = torch.zeros(k1 * k2 * (D + 1), dtype=torch.int64, device="cuda")
    ids0 = torch.full((N,), k1 * k2, dtype=torch.int64, device="cuda")
    ids = kmeans_assign_segmented(torch.from_numpy(a).cuda(), torch.from_numpy(coarse).cuda(), torch.from_numpy(centers).cuda(),
                                  torch.from_numpy(seg_k), k2, ids0, acc, fix)
    assert np.array_equal(ids.cpu().numpy(), want)
    valid = want != k1 * k2
    assert valid.any() or N < 10
    wacc = okm.accumulate_fixed(a, want, k1 * k2, fix, valid)
    assert np.array_equal(acc.cpu().numpy().reshape(k1 * k2, D + 1), wacc)
    # pure reassign (acc = NULL) gives the same ids
    assert torch.equal(kmeans_assign_segmented(torch.from_numpy(a).cuda(), torch.from_numpy(coarse).cuda(),
                                               torch.from_numpy(centers).cuda(), torch.from_numpy(seg_k), k2,
                                               torch.full((N,), k1 * k2, dtype=torch.int64, device="cuda")), ids)


def test_all_leaves_in_one_go_equals_per_cluster_calls():
    """Quantize_kMeans.cluster_assign_all_leaves == the reference's per-cluster leaf calls for every coarse cluster
    (scene/kmeans_quantize.py:196-214,233-238), on the golden scene: same ids away from near-ties, same centres."""
    from opengaussian_b200.kmeans_quantize import Quantize_kMeans
    gold = np.load(GOLD)
    (ins_feat, xyz), (N, k1, k2, iters, pw) = _golden_inputs("root_25k")
    sub = torch.full((k1,), k2, dtype=torch.int64)
    sub[7] = 4
    sub[2] = 1
    qs = []
    for all_at_once in (False, True):
        g = _G()
        g._ins_feat = torch.from_numpy(ins_feat).cuda().requires_grad_(True)
        g._xyz = torch.from_numpy(xyz).cuda()
        q = Quantize_kMeans(num_clusters=k1, num_leaf_clusters=k2, num_iters=iters, dim=9)
        q.cls_ids = torch.from_numpy(gold["root_25k/cls_ids"].astype(np.int64)).cuda()
        q.leaf_centers = torch.from_numpy(ins_feat[:k1 * k2 + 1].copy()).cuda()
        q.leaf_cls_ids = torch.ones(N, device="cuda").to(torch.int64) * k1 * k2
        q.iLeafSubNum = sub
        if all_at_once:
            q.cluster_assign_all_leaves(g)
        else:
            for c in range(k1):
                q.forward(g, 1, assign=True, mode="leaf", selected_leaf=c)
        qs.append((q, g))
    (q0, g0), (q1, g1) = qs
    a, b = q0.leaf_cls_ids.cpu().numpy(), q1.leaf_cls_ids.cpu().numpy()
    assert (a != b).mean() <= 1e-4
    assert a.max() < k1 * k2                                              # every point got a fine id
    assert np.allclose(q0.leaf_centers.cpu().numpy()[:k1 * k2], q1.leaf_centers.cpu().numpy()[:k1 * k2], rtol=1e-4, atol=1e-5)
    assert np.all(q1.leaf_centers.cpu().numpy()[7 * k2 + 4:8 * k2] == 0.0)  # rows beyond iLeafSubNum are rewritten to 0 (:211)
    for sel in (3, 7):                                                    # the two clusters the golden file holds
        m = gold["root_25k/cls_ids"] == sel
        assert (b[m] != gold["root_25k/leaf_cls_ids"].astype(np.int64)[m]).mean() <= 2e-3
    assert torch.allclose(g1._ins_feat_q, q1.leaf_centers[q1.leaf_cls_ids])
    # determinism: exact integer sums -> the same centres bit for bit on a second run
    q2, g2 = Quantize_kMeans(num_clusters=k1, num_leaf_clusters=k2, num_iters=iters, dim=9), _G()
    g2._ins_feat = g1._ins_feat
    q2.cls_ids, q2.iLeafSubNum = q1.cls_ids, sub
    q2.leaf_centers = torch.from_numpy(ins_feat[:k1 * k2 + 1].copy()).cuda()
    q2.cluster_assign_all_leaves(g2)
    assert torch.equal(q2.leaf_centers, q1.leaf_centers) and torch.equal(q2.leaf_cls_ids, q1.leaf_cls_ids)


def test_segmented_full_size_properties():
    """BASELINE config 5's fine level at 5 M points (k1 = 64, k2 = 10, D = 6): every id lies in its coarse cluster's
    block and is the nearest row of that block; counts add up to N."""
    from opengaussian_b200.kmeans_quantize import kmeans_assign_segmented
    N, k1, k2 = 5_000_000, 64, 10
    gen = torch.Generator(device="cuda").manual_seed(1)
    a = torch.rand(N, 6, device="cuda", generator=gen)
    coarse = torch.randint(0, k1, (N,), device="cuda", generator=gen)
    centers = a[:k1 * k2 + 1].contiguous()
    acc = torch.zeros(k1 * k2 * 7, dtype=torch.int64, device="cuda")
    ids = kmeans_assign_segmented(a, coarse, centers, torch.full((k1,), k2, dtype=torch.int32), k2, None, acc, 30)
    assert torch.equal(ids // k2, coarse)
    accv = acc.view(k1 * k2, 7)
    assert int(accv[:, 6].sum()) == N and torch.equal(torch.bincount(ids, minlength=k1 * k2), accv[:, 6])
    idx = torch.randint(0, N, (100_000,), device="cuda", generator=gen)
    blk = centers[:k1 * k2].view(k1, k2, 6)[coarse[idx]].double()         # [n, k2, 6]
    d = (blk - a[idx].double()[:, None, :]).norm(dim=2)
    mine = d.gather(1, (ids[idx] % k2)[:, None])[:, 0]
    assert float((mine - d.min(1).values).max()) <= 1e-5
    want = torch.zeros(k1 * k2, 6, device="cuda", dtype=torch.float64).index_add_(0, ids, a.double())
    assert torch.allclose(accv[:, :6].double() / 2 ** 30, want, rtol=0, atol=N * 2.0 ** -31)


def test_fused_lloyd_pass_equals_separate_kernels():
    """ogs_kmeans_lloyd_pass (assign + sums + centre update in ONE launch, the last CTA finishing the pass) against the
    separate assign / finalize calls: same ids bit for bit, same centres up to the order of the float sums; the
    workspace is reusable across consecutive passes without re-zeroing."""
    import ctypes as C
    from opengaussian_b200 import _lib
    from opengaussian_b200.kmeans_quantize import kmeans_assign
    L = _lib.lib()
    rs = np.random.RandomState(3)
    N, k = 300_001, 64
    a = torch.from_numpy(rs.rand(N, 6).astype(np.float32)).cuda()
    b = torch.from_numpy((rs.rand(N, 3).astype(np.float32) - 0.5) * 6).cuda()
    c0 = torch.cat([a[:k], b[:k] * 0.5], 1).contiguous()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ws = torch.zeros((L.ogs_kmeans_lloyd_workspace_bytes(k, 9) + 3) // 4, dtype=torch.int32, device="cuda")
    cen = c0.clone()
    state = torch.full((k,), 1e-6, device="cuda")
    ids = torch.empty(N, dtype=torch.int64, device="cuda")
    ref_c, ref_state = c0.clone(), torch.full((k,), 1e-6, device="cuda")
    for it in range(3):
        # reference: separate kernels + the Python update of cluster_assign
        sums, cnt = torch.zeros(k, 9, device="cuda"), torch.zeros(k, device="cuda")
        ref_ids = kmeans_assign(a, b, 0.5, ref_c, sums=sums, counts=cnt)
        ref_state += cnt + 31 * 1e-6
        ref_c = sums / ref_state.unsqueeze(-1)
        ref_state[ref_state > 0.1] = 0.
        before = cen.clone()
        rc = L.ogs_kmeans_lloyd_pass(N, a.data_ptr(), 6, b.data_ptr(), 3, 0.5, cen.data_ptr(), k, k, None, -1, 0, ids.data_ptr(),
                                     state.data_ptr(), 31 * 1e-6, None, ws.data_ptr(), stream)
        _lib.check(rc, "ogs_kmeans_lloyd_pass")
        torch.cuda.synchronize()
        assert int(ws[0]) == 0                                            # the ticket was reset by the last CTA
        if it == 0:
            assert torch.equal(ids, ref_ids)                              # same centres in -> same ids out
        else:
            assert torch.equal(ids, kmeans_assign(a, b, 0.5, before))
            assert float((ids != ref_ids).float().mean()) <= 1e-4
        assert torch.allclose(cen, ref_c, rtol=1e-5, atol=1e-6)
        assert torch.allclose(state, ref_state, rtol=1e-5, atol=1e-7)


def test_peer_comm_single_rank_and_empty_shard():
    """The NVLink peer all-reduce with ONE rank is the identity (push into the own inbox, wait on the own flag, sum one
    slot) -- this runs the very kernel the multi-GPU path uses; and a Lloyd pass over an EMPTY shard still takes part
    in the collective and applies the centre update (0 members -> centre 0 / eps)."""
    import ctypes as C
    from opengaussian_b200 import _lib
    L = _lib.lib()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    comm, handle = C.c_void_p(), (C.c_char * 64)()
    _lib.check(L.ogs_peer_comm_create(0, 1, 4480 * 8 + 256, C.byref(comm), handle), "ogs_peer_comm_create")
    _lib.check(L.ogs_peer_comm_connect(comm, bytes(handle)), "ogs_peer_comm_connect")
    try:
        x = torch.randn(640, device="cuda")
        y = torch.randint(-2 ** 40, 2 ** 40, (4480,), device="cuda")
        x0, y0 = x.clone(), y.clone()
        for _ in range(5):                                   # consecutive calls alternate the inbox parity
            _lib.check(L.ogs_peer_allreduce(comm, x.data_ptr(), x.numel(), 0, stream), "ogs_peer_allreduce")
            _lib.check(L.ogs_peer_allreduce(comm, y.data_ptr(), y.numel(), 1, stream), "ogs_peer_allreduce")
        assert L.ogs_peer_comm_error(comm, stream) == 0
        assert torch.equal(x, x0) and torch.equal(y, y0)
        assert L.ogs_peer_allreduce(comm, x.data_ptr(), 1 << 20, 0, stream) != 0      # larger than the slots: refused
        # empty shard through the fused pass, with the communicator in the loop
        k, D = 8, 9
        cen = torch.rand(k, D, device="cuda")
        state = torch.full((k,), 1e-6, device="cuda")
        ws = torch.zeros((L.ogs_kmeans_lloyd_workspace_bytes(k, D) + 3) // 4, dtype=torch.int32, device="cuda")
        rc = L.ogs_kmeans_lloyd_pass(0, None, 6, None, 3, 1.0, cen.data_ptr(), k, k, None, -1, 0, None, state.data_ptr(),
                                     2e-6, comm, ws.data_ptr(), stream)
        _lib.check(rc, "ogs_kmeans_lloyd_pass")
        torch.cuda.synchronize()
        assert float(cen.abs().max()) == 0.0 and torch.allclose(state, torch.full((k,), 3e-6, device="cuda"))
        assert L.ogs_peer_comm_error(comm, stream) == 0
    finally:
        L.ogs_peer_comm_destroy(comm)
