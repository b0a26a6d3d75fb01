"""Developer timing probe (not the judged bench): fwd / bwd ms of a named synthetic scene."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opengaussian_b200 import _lib, synth  # noqa: E402
from opengaussian_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="lerf_1m_1080p")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--P", type=int, default=None)
    ap.add_argument("--fused", type=int, default=0)
    ap.add_argument("--feat_only", type=int, default=0)
    ap.add_argument("--sync", type=int, default=1)
    ap.add_argument("--prof", type=int, default=1)
    a = ap.parse_args()
    gs, cams = synth.make_scene(a.scene, n_views=4, P=a.P)
    dev = "cuda"
    g = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in gs.items()}
    P = g["means3D"].shape[0]
    for vi, cam in enumerate(cams[:2]):
        cam = cam.to(dev)
        rs = GaussianRasterizationSettings(cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy,
                                           torch.zeros(3, device=dev), 1.0, cam.world_view_transform,
                                           cam.full_proj_transform, 3, cam.camera_center, False, False)
        H, W = cam.image_height, cam.image_width
        geom = not a.feat_only
        leaves = dict(means3D=g["means3D"].clone().requires_grad_(geom), opacities=g["opacities"].clone().requires_grad_(geom),
                      shs=g["shs"].clone().requires_grad_(geom), scales=g["scales"].clone().requires_grad_(geom),
                      rotations=g["rotations"].clone().requires_grad_(geom))
        m2 = torch.zeros(P, 3, device=dev, requires_grad=geom)
        extra = g["ins_feat"].clone().requires_grad_(True) if a.fused else None
        gc = torch.randn(3, H, W, device=dev)
        gd = torch.randn(1, H, W, device=dev)
        ga = torch.randn(1, H, W, device=dev)
        gf = torch.randn(6, H, W, device=dev)
        rast = GaussianRasterizer(rs)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tf = tb = 0.0
        hf = hb = 0.0
        t_all0 = None
        for it in range(a.iters + 3):
            for t in list(leaves.values()) + [m2] + ([extra] if extra is not None else []):
                t.grad = None
            if it == 3:
                torch.cuda.synchronize(); t_all0 = time.perf_counter()
                if a.prof:
                    _lib.profile_enable(True); _lib.profile_read()
            ev[0].record()
            h0 = time.perf_counter()
            out = rast(means2D=m2, extra_feats=extra, **leaves)
            h1 = time.perf_counter()
            loss = (out[0] * gc).sum() + (out[2] * gd).sum() + (out[3] * ga).sum()
            if a.fused:
                loss = loss + (out[4] * gf).sum()
            ev[1].record()
            h2 = time.perf_counter()
            loss.backward()
            h3 = time.perf_counter()
            ev[2].record()
            if a.sync:
                torch.cuda.synchronize()
            if it >= 3:
                hf += h1 - h0; hb += h3 - h2
            if it >= 3 and a.sync:
                tf += ev[0].elapsed_time(ev[1])
                tb += ev[1].elapsed_time(ev[2])
        torch.cuda.synchronize()
        if a.prof:
            pr = _lib.profile_read(); _lib.profile_enable(False)
            print("  stages:", {k: round(ms / max(n, 1), 3) for k, (ms, n) in pr.items() if n})
        wall = (time.perf_counter() - t_all0) / a.iters * 1e3
        print(f"  host: fwd call {hf / a.iters * 1e3:.3f} ms, bwd call {hb / a.iters * 1e3:.3f} ms, wall/step {wall:.3f} ms")
        N = out[0].grad_fn.num_rendered if hasattr(out[0].grad_fn, "num_rendered") else -1
        vis = int((out[1] > 0).sum())
        print(f"view {vi}: P={P} vis={vis} N={N} {W}x{H} fwd {tf / a.iters:.3f} ms  bwd {tb / a.iters:.3f} ms  "
              f"fps {1000.0 / ((tf + tb) / a.iters):.1f}", flush=True)


if __name__ == "__main__":
    main()
