"""A few pipelined Stage-1 training steps (BASELINE config 3) exactly as bench.py's stage1 leg issues them -- the
command profiled by `scripts/gpu_evidence_r2.sh r2 c` (ncu --set full over one step)."""
import argparse
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opengaussian_b200 import dist as ogd, rasterizer as rz, synth  # noqa: E402
from opengaussian_b200.mask_stats import cohesion_loss, get_SAM_mask_and_feat, mask_feature_mean, separation_loss  # noqa: E402
from opengaussian_b200.renderer import render  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=6)
ap.add_argument("--view-cache", type=int, default=0, help="1: keep every view's geometry + tile lists resident (rasterizer.ViewCache)")
ap.add_argument("--graph", type=int, default=0, help="1: one CUDA graph per camera (graphs.GraphedViewStep; needs --view-cache 1)")
a = ap.parse_args()
rz.view_cache.enabled = bool(a.view_cache)
dev = torch.device("cuda")
gs, cams = synth.make_scene("scannet_1m_1296x968", n_views=4)
pc = synth.SynthModel(gs, dev, stage0=False)
pipe = types.SimpleNamespace(debug=False, compute_cov3D_python=False, convert_SHs_python=False)
cam_ns = [types.SimpleNamespace(FoVx=c.FoVx, FoVy=c.FoVy, image_height=c.image_height, image_width=c.image_width,
                                world_view_transform=c.world_view_transform.to(dev), full_proj_transform=c.full_proj_transform.to(dev),
                                camera_center=c.camera_center.to(dev), bClusterOccur=None) for c in cams]
H, W = cams[0].image_height, cams[0].image_width
sam_maps = [synth.sam_like_id_map(120, H, W, 4 + v).to(dev) for v in range(4)]
bg = torch.zeros(3, device=dev)


def view_loss(i):
    out = render(cam_ns[i % 4], pc, pipe, bg, 1000, rescale=False)
    _, masks, _ = get_SAM_mask_and_feat(sam_maps[i % 4], level=0, num_mask=120)
    mean = mask_feature_mean(out["ins_feat"], masks, image_mask=out["silhouette"])
    return separation_loss(mean, 1000) + 0.1 * cohesion_loss(out["ins_feat"], masks, mean)


gstep = None
if a.graph:
    from opengaussian_b200.graphs import GraphedViewStep, geometry_guard  # noqa: E402
    gstep = GraphedViewStep(view_loss, [pc._ins_feat], guard=geometry_guard(pc), key=lambda i: i % 4)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(a.iters):
    if it == a.iters // 2:
        torch.cuda.synchronize()
        e0.record()
    if gstep is not None:
        gstep(it)
        continue
    pc._ins_feat.grad = None
    ogd.render_views_backward(view_loss, [it], [pc._ins_feat], already_split=True)
e1.record()
torch.cuda.synchronize()
print(f"{e0.elapsed_time(e1) / (a.iters - a.iters // 2):.3f} ms/step  view_cache={a.view_cache} {rz.view_cache.stats()} graph={gstep.stats() if gstep else None}")
