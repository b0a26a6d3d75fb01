"""Developer timing probe: one training-like render() step (forward + backward to the PARAMETERS) through
opengaussian_b200.renderer.render, raw-parameter path vs the getters' outputs."""
import argparse
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from opengaussian_b200 import synth  # noqa: E402
from opengaussian_b200.renderer import render  # noqa: E402
from test_render_gpu import FakeGaussiansRaw, _cam  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="lerf_1m_1080p")
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    dev = "cuda"
    gs, cams = synth.make_scene(a.scene, n_views=4)
    cam = _cam(cams[1], dev)
    bg = torch.zeros(3, device=dev)
    H, W = cam.image_height, cam.image_width
    G = {k: torch.randn(c, H, W, device=dev) for k, c in (("render", 3), ("ins_feat", 6))}
    for stage, geom_grad in (("stage 0 (all parameters train)", True), ("stage 1 (only ins_feat trains)", False)):
        for raw in (False, True):
            pc = FakeGaussiansRaw(gs, dev, geom_grad=geom_grad)
            pipe = types.SimpleNamespace(debug=False, compute_cov3D_python=False, convert_SHs_python=False,
                                         raw_parameter_path=raw)
            params = [pc._xyz, pc._scaling, pc._rotation, pc._opacity, pc._features_dc, pc._features_rest, pc._ins_feat]

            def step():
                for p in params:
                    p.grad = None
                out = render(cam, pc, pipe, bg, 100, rescale=False)
                torch.autograd.backward((out["render"], out["ins_feat"]), (G["render"], G["ins_feat"]))

            for _ in range(3):
                step()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                step()
            e1.record()
            torch.cuda.synchronize()
            print(f"{a.scene} {stage}: {'raw parameters' if raw else 'getters       '} {e0.elapsed_time(e1) / a.iters:.3f} ms/step")


if __name__ == "__main__":
    main()
