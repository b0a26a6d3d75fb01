"""Developer timing probe for the two section-8f kernels added last: calculate_iou at the Stage-3 association size
(120 SAM masks x 10 leaf silhouettes, 1296x968) and the one-launch Adam step at 1 M Gaussians (65 floats each: the seven groups of training_setup),
beside the reference formulation in plain torch on the same GPU.  Prints one JSON line per probe."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from opengaussian_b200 import synth  # noqa: E402
from opengaussian_b200.mask_stats import calculate_iou  # noqa: E402
from opengaussian_b200.optim import FusedAdam  # noqa: E402
from test_optim_cpu import GROUPS  # noqa: E402


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def reference_iou(masks1, masks2):      # the formulation of utils/opengs_utlis.py:90-123, written out
    a, b = masks1.unsqueeze(0), masks2.unsqueeze(1)
    inter = (a & b).float().sum(dim=(2, 3))
    union = (a | b).float().sum(dim=(2, 3)) + 1e-6
    return inter / union


def main():
    dev = "cuda"
    H, W, n, m = 968, 1296, 120, 10
    m1 = synth.sam_like_masks(n, H, W, 5).to(dev)
    m2 = torch.roll(synth.sam_like_masks(m, H, W, 6).to(dev), (7, -11), dims=(1, 2))
    assert torch.equal(calculate_iou(m1, m2), reference_iou(m1, m2))
    t_mine = timed(lambda: calculate_iou(m1, m2))
    t_ref = timed(lambda: reference_iou(m1, m2), n=5, warm=1)
    print(json.dumps({"probe": "calculate_iou", "masks": [n, m], "image": [W, H], "ms": round(t_mine, 4),
                      "torch_reference_ms": round(t_ref, 3), "mask_bytes_GBps": round((n + m) * H * W / t_mine / 1e6, 1)}))

    P = 1_000_000
    def params():
        g = torch.Generator().manual_seed(0)
        return [torch.nn.Parameter(torch.randn(P, *shape, generator=g).to(dev)) for _, shape, _ in GROUPS]
    def groups(ps):
        return [{"params": [p], "lr": lr, "name": nme} for p, (nme, _, lr) in zip(ps, GROUPS)]
    res = {}
    for name, make in (("fused_one_launch", lambda ps: FusedAdam(groups(ps), lr=0.0, eps=1e-15)),
                       ("torch_default", lambda ps: torch.optim.Adam(groups(ps), lr=0.0, eps=1e-15)),
                       ("torch_fused", lambda ps: torch.optim.Adam(groups(ps), lr=0.0, eps=1e-15, fused=True))):
        ps = params()
        for p in ps:
            p.grad = torch.randn_like(p)
        opt = make(ps)
        res[name] = round(timed(opt.step), 4)
        del opt, ps
        torch.cuda.empty_cache()
    per = sum(int(torch.tensor(shape).prod()) for _, shape, _ in GROUPS)
    elems = P * per
    print(json.dumps({"probe": "adam_step", "gaussians": P, "floats_per_gaussian": per, "ms": res,
                      "algorithmic_GBps": round(elems * 28 / res["fused_one_launch"] / 1e6, 1)}))


if __name__ == "__main__":
    main()
