"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in opengaussian_b200/dist.py and the
sharded mode of Quantize_kMeans.  The CUDA kernels cannot run here, so the k-means compute call is
replaced BY THE TEST with the CPU oracle (the checker); what is under test is the sharding, the
packed [sums|counts] all-reduce per Lloyd iteration, the rank-0 centre broadcast and the gradient
all-reduce -- single-process results are the reference."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_assign(a, b, scale_b, centers, select_ids=None, selected=-1, id_offset=0, ids_out=None, sums=None,
                   counts=None):
    from oracle import kmeans as okm
    an = a.detach().numpy()
    bn = None if b is None else b.detach().numpy()
    sel = None if select_ids is None else select_ids.numpy()
    out = None if ids_out is None else ids_out.numpy()
    k = centers.shape[0]
    ids = okm.assign(an, bn, scale_b, centers.detach().numpy(), sel, selected, id_offset, out)
    if sums is not None or counts is not None:
        s, c = okm.accumulate(an, bn, scale_b, k, ids, sel, selected, id_offset)
        if sums is not None:
            sums += torch.from_numpy(s)
        if counts is not None:
            counts += torch.from_numpy(c)
    return torch.from_numpy(ids) if ids_out is None else ids_out


class _G:
    pass


def _data(N=6000, seed=0):
    rs = np.random.RandomState(seed)
    blobs = rs.rand(12, 6).astype(np.float32)
    feat = (blobs[rs.randint(0, 12, N)] + 0.04 * rs.randn(N, 6)).astype(np.float32)
    xyz = ((rs.rand(N, 3) - 0.5) * 4).astype(np.float32)
    return feat, xyz


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import opengaussian_b200.kmeans_quantize as kq
        from opengaussian_b200 import dist as ogd
        kq.kmeans_assign = _oracle_assign                     # checker stands in for the CUDA call
        kq._GatherStraightThrough.forward = staticmethod(lambda ctx, f, c, ids: c[ids][:, :f.shape[1]].clone())
        feat, xyz = _data()
        lo, hi = ogd.shard_range(feat.shape[0])
        assert (lo, hi) == ((0, 3000) if rank == 0 else (3000, 6000))
        g = _G()
        g._ins_feat = torch.from_numpy(feat[lo:hi]).requires_grad_(True)
        g._xyz = torch.from_numpy(xyz[lo:hi])
        q = ogd.shard_kmeans(kq.Quantize_kMeans(num_clusters=8, num_leaf_clusters=3, num_iters=4, dim=9))
        assert q.distributed
        q.centers = torch.from_numpy(np.concatenate([feat, xyz * np.float32(0.5)], 1)[:8].copy())
        q.forward(g, 1, assign=True, mode="root", pos_weight=0.5)
        ids_all = ogd.gather_ids(q.cls_ids)
        # random-init path: every rank must end with rank 0's draw
        q2 = ogd.shard_kmeans(kq.Quantize_kMeans(num_clusters=5, num_leaf_clusters=3, num_iters=1, dim=9))
        torch.manual_seed(100 + rank)
        q2.forward(g, 1, assign=True, mode="root", pos_weight=0.5)
        c2 = [torch.zeros_like(q2.centers) for _ in range(world)]
        dist.all_gather(c2, q2.centers)

        # gradient all-reduce + view split
        views = list(range(5))
        assert ogd.split_views(views) == views[rank::world]
        p1 = torch.nn.Parameter(torch.arange(6, dtype=torch.float32).reshape(2, 3))
        p2 = torch.nn.Parameter(torch.ones(4))
        p3 = torch.nn.Parameter(torch.ones(2))      # gets no gradient on rank 1

        def render_loss(v):
            loss = (p1 * (v + 1)).sum() + (p2 ** 2).sum() * v
            if rank == 0:
                loss = loss + p3.sum()
            return loss

        total = ogd.render_views_backward(render_loss, views, [p1, p2, p3])

        # fewer views than ranks: rank 1 renders nothing, yet must take part in the same collectives
        # (zero gradients, zero loss) -- a rank-dependent sequence here hangs NCCL
        q1 = torch.nn.Parameter(torch.full((70,), 2.0))
        q2_ = torch.nn.Parameter(torch.ones(3, 3))
        total1 = ogd.render_views_backward(lambda v: (q1 ** 2).sum() + 3.0 * q2_.sum(), [0], [q1, q2_])
        assert q1.grad is not None and q2_.grad is not None and total1 is not None
        few = dict(g1=q1.grad.clone(), g2=q2_.grad.clone(), total=float(total1))
        few_all = [None] * world
        dist.all_gather_object(few_all, few)

        # forward-only sweep: each rank fills the match_info columns of its views, garbage elsewhere
        V = 5
        assert ogd.view_indices(V) == list(range(rank, V, world))
        match_info = torch.full((6, V, 3), float("nan") if rank else -7.0)
        for v in ogd.view_indices(V):
            match_info[:, v] = torch.arange(18, dtype=torch.float32).view(6, 3) + 100 * v
        merged = ogd.merge_view_columns(match_info)
        sub_num = torch.tensor([1, 4, 2], dtype=torch.int32) if rank == 0 else torch.tensor([3, 1, 2], dtype=torch.int32)
        ogd.allreduce_max(sub_num)
        if rank == 0:
            out.put(dict(merged=merged.numpy(), sub_num=sub_num.numpy(),centers=q.centers.numpy(), ids=ids_all.numpy(), c2=[c.numpy() for c in c2],
                         g1=p1.grad.numpy(), g2=p2.grad.numpy(), g3=p3.grad.numpy(), total=float(total),
                         few=[{k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in f.items()} for f in few_all]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_kmeans_and_grad_allreduce_world2():
    from oracle import kmeans as okm
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    feat, xyz = _data()
    c0 = np.concatenate([feat, xyz * np.float32(0.5)], 1)[:8]
    centers, ids = okm.cluster_assign_root(feat, xyz, 0.5, c0, 4)        # single-process reference
    assert (res["ids"] != ids).mean() <= 1e-3
    assert np.allclose(res["centers"], centers, rtol=1e-4, atol=1e-5)
    assert np.array_equal(res["c2"][0], res["c2"][1])                    # broadcast of rank 0's init
    # gradients: sum over all 5 views, as one process rendering every view would get
    v = np.arange(5)
    assert np.allclose(res["g1"], np.full((2, 3), (v + 1).sum(), np.float32))
    assert np.allclose(res["g2"], np.full(4, 2.0 * v.sum(), np.float32))
    assert np.allclose(res["g3"], np.full(2, 3.0, np.float32))           # rank 0 rendered views 0, 2, 4
    w_sum = sum((np.arange(6) * (i + 1)).sum() + 4 * i for i in v) + 3 * 2
    assert abs(res["total"] - w_sum) < 1e-3
    for f in res["few"]:                                                 # both ranks hold the one view's gradients
        assert np.allclose(f["g1"], np.full(70, 4.0, np.float32)) and np.allclose(f["g2"], np.full((3, 3), 3.0, np.float32))
        assert abs(f["total"] - (70 * 4.0 + 27.0)) < 1e-3
    want = np.stack([np.arange(18, dtype=np.float32).reshape(6, 3) + 100 * i for i in range(5)], axis=1)
    assert np.array_equal(res["merged"], want)                           # every view's column, from its owner
    assert res["sub_num"].tolist() == [3, 4, 2]


def test_coalesced_gradient_view_cpu():
    """dist._coalesced_grads: gradients that are 64-float aligned views into one buffer (what the rasterizer's
    backward hands to autograd) are recognised and aliased by ONE flat tensor; anything else is refused."""
    import torch
    from opengaussian_b200 import dist as ogdist
    flat = torch.arange(64 * 5, dtype=torch.float32)
    a, b, c = torch.zeros(10, 3), torch.zeros(7, 4), torch.zeros(5)
    for t in (a, b, c):
        t.requires_grad_(True)
    a.grad = flat[0:30].view(10, 3)
    b.grad = flat[64:92].view(7, 4)
    c.grad = flat[128:133]
    v = ogdist._coalesced_grads([a, b, c])
    assert v is not None and v.numel() == 133
    v.mul_(2.0)
    assert float(a.grad[0, 1]) == 2.0 and float(b.grad[0, 0]) == 128.0 and float(c.grad[4]) == 264.0
    c.grad = torch.zeros(5)                       # a gradient living elsewhere: no coalescing
    assert ogdist._coalesced_grads([a, b, c]) is None
    c.grad = None
    assert ogdist._coalesced_grads([a, b, c]) is None
