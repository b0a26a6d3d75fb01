"""Developer probe: where a Stage-1 training step (BASELINE config 3) spends its time -- device time per phase
(CUDA events) and host time per phase (perf_counter), plus the kernel launch count of one step."""
import os
import sys
import time
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opengaussian_b200 import synth  # noqa: E402
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
sam_maps = [synth.sam_like_id_map(120, H, W, 4 + v).to(dev) for v in range(4)]      # view.original_sam_mask.cuda()
bg = torch.zeros(3, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
names = ["render fwd", "SAM masks", "mask mean", "losses", "backward"]
NP = len(names)
dev_ms = [0.0] * NP
host_ms = [0.0] * NP
n = 30
for it in range(n + 5):
    pc._ins_feat.grad = None
    t = [time.perf_counter()]
    ev[0].record()
    out = render(cam_ns[it % 4], pc, pipe, bg, 1000, rescale=False)
    ev[1].record(); t.append(time.perf_counter())
    _, masks, _ = get_SAM_mask_and_feat(sam_maps[it % 4], level=0, num_mask=120)     # train.py:441
    ev[2].record(); t.append(time.perf_counter())
    mean = mask_feature_mean(out["ins_feat"], masks, image_mask=out["silhouette"])
    ev[3].record(); t.append(time.perf_counter())
    loss = separation_loss(mean, 1000) + 0.1 * cohesion_loss(out["ins_feat"], masks, mean)
    ev[4].record(); t.append(time.perf_counter())
    loss.backward()
    ev[5].record(); t.append(time.perf_counter())
    torch.cuda.synchronize()
    if it >= 5:
        for k in range(NP):
            dev_ms[k] += ev[k].elapsed_time(ev[k + 1]) / n
            host_ms[k] += (t[k + 1] - t[k]) * 1e3 / n
# the same step the way bench.py runs it: no per-step synchronize, through render_views_backward
from opengaussian_b200 import dist as ogd  # noqa: E402


def view_loss(i):
    out = render(cam_ns[i % 4], pc, pipe, bg, 1000, rescale=False)
    _, masks, _ = get_SAM_mask_and_feat(sam_maps[i % 4], level=0, num_mask=120)
    mean = mask_feature_mean(out["ins_feat"], masks, image_mask=out["silhouette"])
    return separation_loss(mean, 1000) + 0.1 * cohesion_loss(out["ins_feat"], masks, mean)


for defer in (False, True):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(3):
        torch.cuda.synchronize()
        e0.record()
        t0 = time.perf_counter()
        for it in range(40):
            pc._ins_feat.grad = None
            ogd.render_views_backward(view_loss, [it], [pc._ins_feat], already_split=True, defer_capacity_check=defer)
        t1 = time.perf_counter()
        e1.record()
        torch.cuda.synchronize()
    print(f"pipelined step, deferred capacity check {defer}: {e0.elapsed_time(e1) / 40:.3f} ms/step (host issue {1e3 * (t1 - t0) / 40:.3f} ms)")
print("phase            device ms   host ms")
for k in range(NP):
    print(f"{names[k]:16s} {dev_ms[k]:9.3f} {host_ms[k]:9.3f}")
print(f"total            {sum(dev_ms):9.3f} {sum(host_ms):9.3f}")
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    pc._ins_feat.grad = None
    out = render(cam_ns[0], pc, pipe, bg, 1000, rescale=False)
    _, masks, _ = get_SAM_mask_and_feat(sam_maps[0], level=0, num_mask=120)
    mean = mask_feature_mean(out["ins_feat"], masks, image_mask=out["silhouette"])
    loss = separation_loss(mean, 1000) + 0.1 * cohesion_loss(out["ins_feat"], masks, mean)
    loss.backward()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
