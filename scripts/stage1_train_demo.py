"""End-to-end Stage-1 loop on the B200 path ("tests/integration"): the instance-feature
training step of train.py:352-456 + :608-610 on a synthetic scene -- ONE fused render (raw parameters), per-mask
feature means, cohesion + separation losses, backward to `_ins_feat`, FusedAdam step -- for N iterations.
Views carry a consistent synthetic "SAM" labelling (every Gaussian belongs to one of K spatial blobs; a view's masks
are the footprints of those blobs), so the loss has something to learn: the run must END with a lower loss than it
started with.  Prints one JSON line (first / last loss averaged over 10 steps, steps/s).  Needs a GPU."""
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
from opengaussian_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer  # noqa: E402
from opengaussian_b200.renderer import render  # noqa: E402


def view_masks(cam, pc, labels, K, dev):
    """[K,H,W] bool: per view, pixel -> blob of the front-most opaque splat (rendered one-hot in chunks of 13 channels)."""
    H, W = cam.image_height, cam.image_width
    rs = GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev), 1.0,
                                       cam.world_view_transform, cam.full_proj_transform, 0, cam.camera_center, False, False)
    rast = GaussianRasterizer(rs)
    onehot = torch.nn.functional.one_hot(labels, K).float()
    maps = []
    with torch.no_grad():
        for c0 in range(0, K, 13):
            extra = onehot[:, c0:c0 + 13].contiguous()
            out = rast(means3D=pc.get_xyz, means2D=torch.zeros_like(pc.get_xyz), opacities=pc.get_opacity,
                       colors_precomp=torch.zeros(labels.shape[0], 3, device=dev), scales=pc.get_scaling,
                       rotations=pc.get_rotation, extra_feats=extra)
            maps.append(out[4])
    votes = torch.cat(maps, 0)                                  # [K,H,W] accumulated weight of every blob
    ids = votes.argmax(0)
    covered = votes.sum(0) > 0.5
    return torch.stack([(ids == k) & covered for k in range(K)])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="blender_300k_800")
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--blobs", type=int, default=24)
    ap.add_argument("--views", type=int, default=8)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    gs, cams = synth.make_scene(a.scene, n_views=a.views)
    pc = synth.SynthModel(gs, dev, stage0=False)
    cams = [c.to(dev) for c in cams]
    g = torch.Generator().manual_seed(0)
    centres = gs["means3D"][torch.randperm(gs["means3D"].shape[0], generator=g)[:a.blobs]]
    labels = torch.cdist(gs["means3D"], centres).argmin(1).to(dev)           # a spatial partition into K blobs
    masks = [view_masks(c, pc, labels, a.blobs, dev) for c in cams]
    masks = [m[m.flatten(1).sum(1) > 50] for m in masks]                     # drop tiny masks (train.py filters likewise)
    pipe = types.SimpleNamespace(debug=False, compute_cov3D_python=False, convert_SHs_python=False)
    bg = torch.zeros(3, device=dev)
    opt = FusedAdam([{"params": [pc._ins_feat], "lr": 0.001, "name": "ins_feat"}], lr=0.0, eps=1e-15)
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
    print(json.dumps({"probe": "stage1_train_demo", "scene": a.scene, "iters": a.iters, "blobs": a.blobs,
                      "masks_per_view": [int(m.shape[0]) for m in masks], "loss_first10": round(first, 5),
                      "loss_last10": round(last, 5), "decreased": last < first, "ms_per_step": round(ms, 3),
                      "steps_per_s": round(1e3 / ms, 1)}))
    assert last < first, "Stage-1 loss did not decrease"


if __name__ == "__main__":
    main()
