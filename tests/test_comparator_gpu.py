"""GPU comparator (baseline/upstream_structure.cu: the upstream-STRUCTURE restatement the ">= 10x the reference
CUDA rasterizer" figure is measured against) checked against the CPU oracle: a comparator that computed something
else would make the speed-up meaningless.  Same bars as the product's parity tests."""
import numpy as np
import pytest
import torch

from helpers import grad_violations, np_inputs, small_scene, to_oracle_cam
from oracle import raster as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("P,W,H,scale_mult", [(500, 80, 64, 2.5), (3000, 96, 96, 4.0)])
def test_upstream_structure_vs_oracle(P, W, H, scale_mult):
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline"))
    import comparator as cmp
    dev = torch.device("cuda", 0)
    gs, cam = small_scene(P=P, W=W, H=H, seed=1, scale_mult=scale_mult)
    g = np_inputs(gs)
    bg = (0.1, 0.2, 0.3)
    st = orc.forward(to_oracle_cam(cam), g["means3D"], g["opacities"], g["scales"], g["rotations"], shs=g["shs"],
                     bg=np.array(bg, np.float32))
    p = cmp.Pass(cmp.entry_points("upstream_structure"), gs, cam, dev, bg=bg)
    ok = st.flags == 0
    m = torch.as_tensor(ok, device=dev).float()
    p.g_color *= m
    p.g_depth *= m
    p.g_alpha *= m
    n = p.run()
    torch.cuda.synchronize()
    assert n == st.N
    assert np.array_equal(p.out["radii"].cpu().numpy(), st.radii)
    assert np.abs(p.out["color"].cpu().numpy() - st.color)[:, ok].max() <= 1e-5
    assert np.abs(p.out["depth"].cpu().numpy() - st.out_depth)[ok].max() <= 1e-5 * max(1.0, float(st.out_depth.max()))
    assert np.abs(p.out["alpha"].cpu().numpy() - st.out_alpha)[ok].max() <= 1e-5
    ref = orc.backward(st, p.g_color.cpu().numpy(), p.g_depth.cpu().numpy(), p.g_alpha.cpu().numpy())
    for k, v in p.grads.items():
        want = np.asarray(ref[k])
        got = v.cpu().numpy().reshape(want.shape)
        frac, worst = grad_violations(got, want)
        print(f"comparator dL/d{k}: violating fraction {frac:.2e}, worst |diff|/tol {worst:.3f}")
        assert frac == 0.0, (k, frac, worst)
