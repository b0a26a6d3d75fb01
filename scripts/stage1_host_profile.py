"""Developer probe: host-side (Python + launch) profile of a Stage-1 training step (BASELINE config 3)."""
import cProfile
import os
import pstats
import sys
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
sam_maps = [synth.sam_like_id_map(120, H, W, 4 + v).to(dev) for v in range(4)]
bg = torch.zeros(3, device=dev)


def step(it):
    pc._ins_feat.grad = None
    out = render(cam_ns[it % 4], pc, pipe, bg, 1000, rescale=False)
    _, masks, _ = get_SAM_mask_and_feat(sam_maps[it % 4], level=0, num_mask=120)
    mean = mask_feature_mean(out["ins_feat"], masks, image_mask=out["silhouette"])
    loss = separation_loss(mean, 1000) + 0.1 * cohesion_loss(out["ins_feat"], masks, mean)
    loss.backward()


for it in range(10):
    step(it)
torch.cuda.synchronize()
N = 200
pr = cProfile.Profile()
pr.enable()
for it in range(N):
    step(it)
    if it % 8 == 7:
        torch.cuda.synchronize()          # keep the launch queue from filling: host time only
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(32)
st.sort_stats("cumtime").print_stats(28)
