"""Developer probe: device timeline of fwd+bwd frames of a synthetic scene from torch.profiler (CUPTI) -- per kernel the
mean duration, and busy time against wall time per frame: what the gaps between a frame's ~20 dependent launches cost
(a launch list taken under ncu serialises the kernels and cannot show it)."""
import argparse
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opengaussian_b200 import rasterizer as rz, synth  # noqa: E402
from opengaussian_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="lerf_1m_1080p")
ap.add_argument("--frames", type=int, default=8)
a = ap.parse_args()
dev = "cuda"
gs, cams = synth.make_scene(a.scene, n_views=4)
g = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in gs.items()}
P = g["means3D"].shape[0]
leaves = {k: g[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
views = []
for cam in cams:
    cam = cam.to(dev)
    rs = GaussianRasterizationSettings(cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev),
                                       1.0, cam.world_view_transform, cam.full_proj_transform, 3, cam.camera_center, False, False)
    H, W = cam.image_height, cam.image_width
    views.append((GaussianRasterizer(rs), torch.randn(3, H, W, device=dev), torch.randn(1, H, W, device=dev),
                  torch.randn(1, H, W, device=dev)))


def frame(i):
    rast, gc, gd, ga = views[i % len(views)]
    m2 = torch.zeros(P, 3, device=dev, requires_grad=True)
    color, radii, depth, alpha = rast(means2D=m2, **leaves)
    torch.autograd.backward([color, depth, alpha], [gc, gd, ga])


for i in range(8):
    frame(i)
torch.cuda.synchronize()
with rz.deferred_capacity_check():
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(a.frames):
            frame(i)
        torch.cuda.synchronize()
    rz.capacity_overflowed(torch.device(dev))
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
acc = collections.OrderedDict()
for e in ev:
    x = acc.setdefault(e.name[:64], [0, 0.0])
    x[0] += 1
    x[1] += e.time_range.end - e.time_range.start
N = a.frames
busy = sum(x[1] for x in acc.values())
wall = ev[-1].time_range.end - ev[0].time_range.start
print(f"{len(ev) / N:.1f} device activities per frame; busy {busy / N:.1f} us, wall {wall / N:.1f} us per frame")
for k, x in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:66s} {x[0] / N:5.1f} x {x[1] / x[0]:7.1f} us = {x[1] / N:7.1f} us/frame")
gaps = [(y.time_range.start - x.time_range.end, x.name[:36], y.name[:36]) for x, y in zip(ev[:-1], ev[1:])]
print(f"sum of positive gaps per frame: {sum(max(gp[0], 0) for gp in gaps) / N:.1f} us")
by = collections.Counter()
for gp in gaps:
    by[(gp[1], gp[2])] += max(gp[0], 0)
print("gaps by kernel pair (us per frame):")
for (x, y), v in by.most_common(14):
    print(f"  {v / N:6.1f}  {x}  ->  {y}")
