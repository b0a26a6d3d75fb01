"""Developer timing probe for the k-means assign (+ fused centroid sums) kernel."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opengaussian_b200 import _lib  # noqa: E402
from opengaussian_b200.kmeans_quantize import kmeans_assign  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=5_000_000)
    ap.add_argument("--k", type=int, default=64)
    ap.add_argument("--fuse", type=int, default=1)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--coherent", type=int, default=0, help="points of one cluster are contiguous in memory (trained scenes are spatially coherent)")
    a = ap.parse_args()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(7)
    fa = torch.rand(a.N, 6, device=dev, generator=g)
    fb = (torch.rand(a.N, 3, device=dev, generator=g) - 0.5) * 8
    cen = torch.cat([fa[:a.k], fb[:a.k]], 1).contiguous()
    if a.coherent:
        own = (torch.arange(a.N, device=dev) * a.k // a.N)
        fa = (cen[own, :6] + 0.01 * torch.randn(a.N, 6, device=dev, generator=g)).contiguous()
        fb = (cen[own, 6:] + 0.01 * torch.randn(a.N, 3, device=dev, generator=g)).contiguous()
    ids = torch.empty(a.N, dtype=torch.int64, device=dev)
    s9 = torch.zeros(a.k, 9, device=dev)
    c1 = torch.zeros(a.k, device=dev)
    for variant, (b_, D) in {"D=9 (ins_feat|xyz)": (fb, 9), "D=6 (ins_feat)": (None, 6)}.items():
        c = cen if b_ is not None else cen[:, :6].contiguous()
        s = s9 if b_ is not None else torch.zeros(a.k, 6, device=dev)
        for _ in range(3):
            kmeans_assign(fa, b_, 1.0, c, ids_out=ids, sums=s if a.fuse else None, counts=c1 if a.fuse else None)
        _lib.profile_enable(True); _lib.profile_read()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            kmeans_assign(fa, b_, 1.0, c, ids_out=ids, sums=s if a.fuse else None, counts=c1 if a.fuse else None)
        e1.record()
        torch.cuda.synchronize()
        pr = _lib.profile_read(); _lib.profile_enable(False)
        ms = e0.elapsed_time(e1) / a.iters
        kms, n = pr["kmeans_assign"]
        print(f"{variant} k={a.k} N={a.N} fuse={a.fuse}: {ms:.4f} ms/pass wall-on-device, kernel family {kms / max(n, 1):.4f} ms "
              f"-> {a.N / (kms / max(n, 1)) / 1e6:.2f} Gpts/s, HBM {a.N * (4 * D + 8) / (kms / max(n, 1) * 1e-3) / 1e9:.0f} GB/s")


def leaf():
    """Leaf-level pass as the reference drives it: one coarse cluster per call (train.py:330-332)."""
    dev = "cuda"
    N, k1, k2 = 5_000_000, 64, 10
    g = torch.Generator(device=dev).manual_seed(9)
    fa = torch.rand(N, 6, device=dev, generator=g)
    coarse = torch.randint(0, k1, (N,), device=dev, generator=g)
    cen = fa[:k2].contiguous()
    ids = torch.zeros(N, dtype=torch.int64, device=dev)
    s = torch.zeros(k2, 6, device=dev); c = torch.zeros(k2, device=dev)
    for _ in range(3):
        kmeans_assign(fa, None, 1.0, cen, select_ids=coarse, selected=3, id_offset=30, ids_out=ids, sums=s, counts=c)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for leaf_id in range(k1):
        kmeans_assign(fa, None, 1.0, cen, select_ids=coarse, selected=leaf_id, id_offset=leaf_id * k2, ids_out=ids, sums=s, counts=c)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"leaf level: 64 per-cluster passes over N={N} (k2={k2}, D=6): {ms:.3f} ms total, {ms / k1 * 1e3:.1f} us/pass, "
          f"{N / (ms * 1e-3) / 1e9:.2f} Gpts/s (every point assigned once)")


if __name__ == "__main__":
    if "--leaf" in sys.argv:
        sys.argv.remove("--leaf")
        leaf()
        sys.exit(0)
    main()
