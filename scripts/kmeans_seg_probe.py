"""Timing probe of the one-launch Lloyd passes (coarse k=64 D=9, fine 64x10 D=6) at several point counts."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opengaussian_b200.kmeans_quantize import LloydWorkspace, kmeans_assign, lloyd_pass, lloyd_pass_segmented  # noqa: E402

dev = torch.device("cuda")
for N in (5_000_000, 2_500_000, 1_250_000, 625_000):
    g = torch.Generator(device=dev).manual_seed(7)
    fa = torch.rand(N, 6, device=dev, generator=g)
    fb = (torch.rand(N, 3, device=dev, generator=g) - 0.5) * 8
    cen = torch.cat([fa[:64], fb[:64]], 1).contiguous()
    ids = torch.empty(N, dtype=torch.int64, device=dev)
    kmeans_assign(fa, fb, 1.0, cen, ids_out=ids)
    coarse = ids.clone()
    leaf_c = fa[:641].contiguous().clone()
    seg_k = torch.full((64,), 10, dtype=torch.int32, device=dev)
    ws_c, ws_f = LloydWorkspace(dev, k=64, D=9), LloydWorkspace(dev, k1=64, k2=10, D=6)
    sc, sf = torch.full((64,), 1e-6, device=dev), torch.full((640,), 1e-6, device=dev)
    lids = torch.empty(N, dtype=torch.int64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    res = {}
    for name, fn in (("coarse", lambda: lloyd_pass(fa, fb, 1.0, cen, sc, 1e-6, ids, ws_c)),
                     ("fine", lambda: lloyd_pass_segmented(fa, coarse, leaf_c, seg_k, 10, sf, lids, 30, ws_f))):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 50 * 1e3
    print(f"N={N}: coarse {res['coarse']:.1f} us ({N / res['coarse'] / 1e3:.1f} Gpts/s)  fine {res['fine']:.1f} us ({N / res['fine'] / 1e3:.1f} Gpts/s)",
          flush=True)
