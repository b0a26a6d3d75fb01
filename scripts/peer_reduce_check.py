"""Multi-GPU check + timing of the one-kernel NVLink all-reduce (csrc/peer.cu, dist.PeerReducer) and of the sharded
two-level k-means on top of it.  Run under torchrun on one node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/peer_reduce_check.py

Checks (asserted on every rank): float32 and int64 results equal NCCL's (int64 exactly; float32 exactly when the
rank-ordered sum is what NCCL computes, else to 1e-6) and are bit-identical across ranks, over 2000 back-to-back calls
of varying length (the two-parity inbox protocol); the sharded fine-level k-means (exact integer sums) reproduces the
single-GPU centres BIT FOR BIT.  Prints one JSON line on rank 0 with the per-call latency of both collectives."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opengaussian_b200 import dist as ogd  # noqa: E402
from opengaussian_b200.kmeans_quantize import Quantize_kMeans  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    red = ogd.PeerReducer(64 * 10 * 8 * 8 + 4096)
    out = {"world": world, "kind": red.kind}
    assert red.comm is not None, "peer memory could not be mapped: " + red.kind
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    # ---- correctness over many back-to-back calls ----
    worst = 0.0
    for it in range(2000):
        n = 1 + (it * 37) % 4480
        if it % 2 == 0:
            x = torch.randn(n, device=dev, generator=g)
            want = x.clone()
            dist.all_reduce(want)
            red.all_reduce(x)
            worst = max(worst, float((x - want).abs().max() / (want.abs().max() + 1e-20)))
        else:
            x = torch.randint(-2 ** 40, 2 ** 40, (n,), device=dev, generator=g)
            want = x.clone()
            dist.all_reduce(want)
            red.all_reduce(x)
            assert torch.equal(x, want), it
        if it % 500 == 0:                    # identical bits on every rank
            xs = [torch.empty_like(x) for _ in range(world)]
            dist.all_gather(xs, x)
            assert all(torch.equal(xs[0], t) for t in xs)
    red.check()
    assert worst <= 1e-6, worst
    out["float_max_rel_diff_vs_nccl"] = worst
    # ---- latency ----
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, t in (("f32_640", torch.ones(640, device=dev)), ("i64_4480", torch.ones(4480, dtype=torch.int64, device=dev))):
        for label, fn in (("peer", lambda: red.all_reduce(t)), ("nccl", lambda: dist.all_reduce(t))):
            for _ in range(20):
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            e0.record()
            for _ in range(200):
                fn()
            e1.record()
            torch.cuda.synchronize()
            us = torch.tensor([e0.elapsed_time(e1) / 200 * 1e3], device=dev)
            dist.all_reduce(us, op=dist.ReduceOp.MAX)
            out[f"{label}_{name}_us"] = round(float(us), 2)
            t.fill_(1)
    # ---- sharded two-level k-means == single GPU, bit for bit at the fine level ----
    N, k1, k2 = 400_000, 16, 5
    gg = torch.Generator(device=dev).manual_seed(5)           # same data on every rank
    blobs = torch.rand(40, 6, device=dev, generator=gg) * 2 - 1   # clustered features: k-means has structure to find
    feat = blobs[torch.randint(0, 40, (N,), device=dev, generator=gg)] + 0.05 * torch.randn(N, 6, device=dev, generator=gg)
    xyz = torch.rand(N, 3, device=dev, generator=gg) * 4

    class G:
        pass

    def run(lo, hi, sharded):
        gm = G()
        gm._ins_feat = feat[lo:hi].clone().requires_grad_(True)
        gm._xyz = xyz[lo:hi].clone()
        q = Quantize_kMeans(num_clusters=k1, num_leaf_clusters=k2, num_iters=4, dim=9)
        if sharded:
            ogd.shard_kmeans(q, peer_reduce=True)
        q.centers = torch.cat([feat[:k1], xyz[:k1] * 0.5], 1).contiguous()
        q.forward(gm, 1, assign=True, mode="root", pos_weight=0.5)
        q.leaf_centers = feat[:k1 * k2 + 1].clone()
        q.iLeafSubNum = torch.full((k1,), k2, dtype=torch.int64)
        q.cluster_assign_all_leaves(gm)
        return q

    lo, hi = ogd.shard_range(N)
    qs = run(lo, hi, True)
    qs.reducer.check()
    q1 = run(0, N, False)
    # coarse level: float sums in a different order -> centres agree to rounding, ids away from near-ties
    centre_diff = float((qs.centers - q1.centers).abs().max() / q1.centers.abs().max())
    coarse_mismatch = float((qs.cls_ids != q1.cls_ids[lo:hi]).float().mean())
    out["coarse_centre_max_rel_diff_vs_single"] = centre_diff
    assert centre_diff <= 2e-3 and coarse_mismatch <= 2e-3, (centre_diff, coarse_mismatch)
    # fine level with identical coarse ids: exact integer sums -> bit-identical centres and ids
    qs2, q12 = Quantize_kMeans(k1, k2, 4, 9), Quantize_kMeans(k1, k2, 4, 9)
    ogd.shard_kmeans(qs2, peer_reduce=True)
    for q, (a, b) in ((qs2, (lo, hi)), (q12, (0, N))):
        q.cls_ids = q1.cls_ids[a:b].clone()
        q.leaf_centers = feat[:k1 * k2 + 1].clone()
        q.iLeafSubNum = torch.full((k1,), k2, dtype=torch.int64)
        q.cluster_assign_all_leaves(feat=feat[a:b])
    assert torch.equal(qs2.leaf_centers, q12.leaf_centers)
    assert torch.equal(qs2.leaf_cls_ids, q12.leaf_cls_ids[lo:hi])
    out["sharded_fine_level_bit_identical"] = True
    out["coarse_id_mismatch_vs_single"] = coarse_mismatch
    qs.reducer.close()
    qs2.reducer.close()
    red.close()
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
