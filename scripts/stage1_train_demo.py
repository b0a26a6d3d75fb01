"""End-to-end Stage-1 loop on the B200 path: the instance-feature training step of train.py:352-456 + :608-610 on a
synthetic scene -- ONE fused render (raw parameters), per-mask feature means, cohesion + separation losses, backward
to `_ins_feat`, FusedAdam step -- for N iterations.  Views carry a consistent synthetic "SAM" labelling (every
Gaussian belongs to one of K spatial blobs; a view's masks are the footprints of those blobs), so the loss has
something to learn: the run must END with a lower loss than it started with.  Prints one JSON line (first / last loss
averaged over 10 steps, steps/s).  Needs a GPU.  (tests/test_integration_gpu.py runs the same loop as a test.)"""
import argparse
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opengaussian_b200 import synth  # noqa: E402
from opengaussian_b200.mask_stats import cohesion_loss, mask_feature_mean, separation_loss  # noqa: E402
from opengaussian_b200.optim import FusedAdam  # noqa: E402
from opengaussian_b200.renderer import render  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="blender_300k_800")
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--blobs", type=int, default=24)
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--lr", type=float, default=0.001, help="ins_feat_lr of the reference (arguments/__init__.py:80)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    gs, cams = synth.make_scene(a.scene, n_views=a.views)
    pc = synth.SynthModel(gs, dev, stage0=False)
    cams = [c.to(dev) for c in cams]
    labels = synth.blob_labels(gs, a.blobs).to(dev)
    masks = [synth.blob_view_masks(c, pc, labels, a.blobs) for c in cams]
    pipe = types.SimpleNamespace(debug=False, compute_cov3D_python=False, convert_SHs_python=False)
    bg = torch.zeros(3, device=dev)
    opt = FusedAdam([{"params": [pc._ins_feat], "lr": a.lr, "name": "ins_feat"}], lr=0.0, eps=1e-15)
    losses = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(a.iters):
        if it == 10:
            e0.record()
        v = it % len(cams)
        out = render(cams[v], pc, pipe, bg, it, rescale=False)
        mean = mask_feature_mean(out["ins_feat"], masks[v], image_mask=out["silhouette"])
        loss = separation_loss(mean, it) + 0.1 * cohesion_loss(out["ins_feat"], masks[v], mean)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        losses.append(loss.detach())
    e1.record()
    torch.cuda.synchronize()
    losses = torch.stack(losses).cpu()
    first, last = float(losses[:10].mean()), float(losses[-10:].mean())
    ms = e0.elapsed_time(e1) / max(1, a.iters - 10)
    print(json.dumps({"probe": "stage1_train_demo", "scene": a.scene, "iters": a.iters, "blobs": a.blobs, "lr": a.lr,
                      "masks_per_view": [int(m.shape[0]) for m in masks], "loss_first10": round(first, 5),
                      "loss_last10": round(last, 5), "decreased": last < first, "ms_per_step": round(ms, 3),
                      "steps_per_s": round(1e3 / ms, 1)}))
    assert last < first, "Stage-1 loss did not decrease"


if __name__ == "__main__":
    main()
