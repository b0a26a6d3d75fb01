"""Developer probe: device timeline of graphed Stage-1 steps (scripts/stage1_steps.py --view-cache 1 --graph 1) from
torch.profiler (CUPTI): per kernel the mean duration and, over the replays, busy time against wall time -- the gaps
between the ~40 nodes of a step are what a launch list taken under ncu cannot show."""
import collections
import os
import sys
import types

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opengaussian_b200 import rasterizer as rz, synth  # noqa: E402
from opengaussian_b200.graphs import GraphedViewStep, geometry_guard  # noqa: E402
from opengaussian_b200.mask_stats import cohesion_loss, get_SAM_mask_and_feat, mask_feature_mean, separation_loss  # noqa: E402
from opengaussian_b200.renderer import render  # noqa: E402

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
rz.view_cache.enabled = True


def view_loss(i):
    out = render(cam_ns[i % 4], pc, pipe, bg, 1000, rescale=False)
    _, masks, _ = get_SAM_mask_and_feat(sam_maps[i % 4], level=0, num_mask=120)
    mean = mask_feature_mean(out["ins_feat"], masks, image_mask=out["silhouette"])
    return separation_loss(mean, 1000) + 0.1 * cohesion_loss(out["ins_feat"], masks, mean)


gstep = GraphedViewStep(view_loss, [pc._ins_feat], guard=geometry_guard(pc), key=lambda i: i % 4)
for it in range(16):
    gstep(it)
torch.cuda.synchronize()
N = 8
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for it in range(N):
        gstep(it)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
acc = collections.OrderedDict()
for e in ev:
    a = acc.setdefault(e.name[:70], [0, 0.0])
    a[0] += 1
    a[1] += e.time_range.end - e.time_range.start
busy = sum(a[1] for a in acc.values())
wall = ev[-1].time_range.end - ev[0].time_range.start
print(f"{len(ev) / N:.1f} device activities per step; busy {busy / N:.1f} us, wall {wall / N:.1f} us per step")
for k, a in sorted(acc.items(), key=lambda x: -x[1][1]):
    print(f"{k:72s} {a[0] / N:5.1f} x {a[1] / a[0]:7.1f} us = {a[1] / N:7.1f} us/step")
# the largest gaps
gaps = []
for x, y in zip(ev[:-1], ev[1:]):
    gaps.append((y.time_range.start - x.time_range.end, x.name[:40], y.name[:40]))
gaps.sort(reverse=True)
print("largest gaps (us, after, before):")
for g in gaps[:12]:
    print(f"  {g[0]:7.1f}  {g[1]}  ->  {g[2]}")
print(f"sum of positive gaps per step: {sum(max(g[0], 0) for g in gaps) / N:.1f} us")
