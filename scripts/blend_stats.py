"""Developer probe: how much of the blend work is useful?  For a sample of tiles of a synthetic
scene, recompute (in torch, from the exported binning state) per (pixel, list entry) whether the
entry contributes, and report the efficiency of culling at different block granularities."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opengaussian_b200 import synth  # noqa: E402
from opengaussian_b200.debug import forward_with_state  # noqa: E402
from opengaussian_b200.rasterizer import GaussianRasterizationSettings  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="lerf_1m_1080p")
    ap.add_argument("--tiles", type=int, default=400)
    a = ap.parse_args()
    gs, cams = synth.make_scene(a.scene, n_views=2)
    dev = "cuda"
    g = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in gs.items()}
    cam = cams[0].to(dev)
    H, W = cam.image_height, cam.image_width
    rs = GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev), 1.0,
                                       cam.world_view_transform, cam.full_proj_transform, 3, cam.camera_center, False, False)
    st = forward_with_state(rs, g["means3D"], g["opacities"], shs=g["shs"], scales=g["scales"], rotations=g["rotations"])
    ranges = st["ranges"].long()
    gx = (W + 15) // 16
    ntiles = ranges.shape[0]
    gen = torch.Generator().manual_seed(0)
    sel = torch.randperm(ntiles, generator=gen)[: a.tiles]
    tot = dict(entries=0, px_entries=0, px_before_stop=0, contrib=0)
    blocks = {"8x8": (8, 8), "8x16": (8, 16), "16x16": (16, 16), "8x4": (8, 4), "4x4": (4, 4)}
    blk_eval = {k: 0 for k in blocks}       # (block, entry) evaluations x pixels in block, with block-level early stop + exact any-contrib culling
    blk_count = {k: 0 for k in blocks}      # (block, entry) evaluations
    for t in sel.tolist():
        lo, hi = ranges[t].tolist()
        L = hi - lo
        if L == 0:
            continue
        ids = st["point_list"][lo:hi].long()
        xy = st["xy"][ids]
        co = st["conic_opacity"][ids]
        tx, ty = t % gx, t // gx
        px = (tx * 16 + torch.arange(16, device=dev)).float()
        py = (ty * 16 + torch.arange(16, device=dev)).float()
        PY, PX = torch.meshgrid(py, px, indexing="ij")
        inside = ((PX < W) & (PY < H)).reshape(-1)
        dx = xy[:, 0][None, :] - PX.reshape(-1, 1)
        dy = xy[:, 1][None, :] - PY.reshape(-1, 1)
        power = -0.5 * (co[:, 0] * dx * dx + co[:, 2] * dy * dy) - co[:, 1] * dx * dy
        alpha = torch.clamp(co[:, 3] * torch.exp(power), max=0.99)
        ok = (power <= 0) & (alpha >= 1.0 / 255.0) & inside[:, None]
        a_eff = torch.where(ok, alpha, torch.zeros_like(alpha))
        Tcum = torch.cumprod(1 - a_eff, dim=1)          # T after applying entry j
        stop = (Tcum < 1e-4) & ok
        # index of first stop (entry not applied), L if none
        first = torch.where(stop.any(1), stop.float().argmax(1), torch.full((256,), L, device=dev))
        idx = torch.arange(L, device=dev)[None, :]
        live = idx < first[:, None]                      # entries the pixel processes before stopping
        contrib = ok & live
        tot["entries"] += L
        tot["px_entries"] += 256 * L
        tot["px_before_stop"] += int(live.sum())
        tot["contrib"] += int(contrib.sum())
        c2 = contrib.reshape(16, 16, L)
        l2 = live.reshape(16, 16, L)
        for name, (bw, bh) in blocks.items():
            cb = c2.reshape(16 // bh, bh, 16 // bw, bw, L).permute(0, 2, 1, 3, 4).reshape(-1, bh * bw, L)
            lb = l2.reshape(16 // bh, bh, 16 // bw, bw, L).permute(0, 2, 1, 3, 4).reshape(-1, bh * bw, L)
            anyc = cb.any(1)                  # block has at least one contributing pixel for the entry
            blk_count[name] += int(anyc.sum())
            blk_eval[name] += int(anyc.sum()) * bw * bh
            del cb, lb
    print(f"tiles sampled {len(sel)}  mean list {tot['entries'] / len(sel):.1f}")
    print(f"pixel-entries total {tot['px_entries']:.3e}  before stop {tot['px_before_stop']:.3e} "
          f"({tot['px_before_stop'] / tot['px_entries']:.3f})  contributing {tot['contrib']:.3e} "
          f"({tot['contrib'] / tot['px_entries']:.3f})")
    for name in blocks:
        print(f"  block {name}: (block,entry) evals {blk_count[name]:.3e}  pixel-evals {blk_eval[name]:.3e}  "
              f"useful fraction {tot['contrib'] / max(blk_eval[name], 1):.3f}")


if __name__ == "__main__":
    main()
