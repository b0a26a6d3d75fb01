"""GPU comparator run (last row): the upstream-STRUCTURE restatement of the reference CUDA
rasterizer (baseline/upstream_structure.cu -- one thread per pixel, 64-bit key library sort, per-pair global
atomics; NOT the reference's code, whose source is absent) against the product, through the same raw C-ABI calls
(baseline/comparator.py):
  1. parity on the plumbing scene: images 2e-5 absolute (expf vs ex2.approx), gradients 1e-3 relative;
  2. fwd+bwd frames/s of both on the BASELINE workload (1 M Gaussians, 1920x1080, SH degree 3, all gradients);
  3. the rasterizer work of one Stage-1 step on the ScanNet-like config: the reference's 4 forward + 2 backward
     3-channel passes (gaussian_renderer/__init__.py:104-163) against the product's single fused pass.
Prints one JSON line.  Needs a GPU; run `python baseline/build_comparator.py` first."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import comparator as cmp  # noqa: E402
from opengaussian_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="lerf_1m_1080p")
    ap.add_argument("--stage1-scene", default="scannet_1m_1296x968")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--only", default="", help="'upstream': run only the comparator's frame (for an ncu launch list)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    if a.only == "upstream":
        gs, cams = synth.make_scene(a.scene, n_views=2)
        p = cmp.Pass(cmp.entry_points("upstream_structure"), gs, cams[0], dev)
        print(json.dumps({"ms_per_frame": cmp.time_passes([p], a.iters, a.warmup, dev)}))
        return

    # 1. parity on the plumbing scene
    gs, cams = synth.make_scene("plumbing_10k_256", n_views=2)
    fa = cmp.Pass(cmp.entry_points("product"), gs, cams[0], dev)
    fb = cmp.Pass(cmp.entry_points("upstream_structure"), gs, cams[0], dev)
    na, nb = fa.run(), fb.run()
    torch.cuda.synchronize()
    parity = {"num_rendered_equal": na == nb, "radii_equal": bool(torch.equal(fa.out["radii"], fb.out["radii"]))}
    for k in ("color", "depth", "alpha"):
        parity[f"{k}_max_abs"] = float((fa.out[k] - fb.out[k]).abs().max())
    for k in fa.grads:
        parity[f"d{k}_rel"] = float((fa.grads[k] - fb.grads[k]).abs().max() / (fa.grads[k].abs().max() + 1e-20))
    ok = parity["num_rendered_equal"] and parity["radii_equal"] and \
        all(parity[f"{k}_max_abs"] <= 2e-5 * max(1.0, float(fa.out[k].abs().max())) for k in ("color", "depth", "alpha")) and \
        all(parity[f"d{k}_rel"] <= 1e-3 for k in fa.grads)
    del fa, fb

    # 2. the BASELINE frame; 3. the Stage-1 step's rasterizer work
    gs, cams = synth.make_scene(a.scene, n_views=2)
    frame = cmp.frame_comparison(gs, cams[0], dev, a.iters, a.warmup)
    gs, cams = synth.make_scene(a.stage1_scene, n_views=2)
    stage1 = cmp.stage1_comparison(gs, cams[0], dev, a.iters, a.warmup)
    print(json.dumps({"probe": "upstream_structure_comparator", "scene": a.scene, "parity_ok": bool(ok), "parity": parity,
                      "frame": frame, "stage1_scene": a.stage1_scene, "stage1": stage1, "note": cmp.NOTE}))


if __name__ == "__main__":
    main()
