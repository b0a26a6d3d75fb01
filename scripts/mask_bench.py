"""Developer timing probe: Stage-1 loss (mask_feature_mean + cohesion + separation, fwd + bwd) at the
BASELINE config-3 image size, B200 kernels vs the reference formulation in plain torch on the GPU."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from opengaussian_b200 import _lib  # noqa: E402
from opengaussian_b200.mask_stats import cohesion_loss, mask_feature_mean, separation_loss  # noqa: E402
from test_mask_stats_gpu import _sam_like_masks  # noqa: E402


def main():
    C, H, W, M = 6, 968, 1296, 120
    dev = "cuda"
    g = torch.Generator().manual_seed(3)
    feat = torch.rand(C, H, W, generator=g).to(dev).requires_grad_(True)
    img = ((torch.rand(1, H, W, generator=g) > 0.1).float() * torch.rand(1, H, W, generator=g)).to(dev).requires_grad_(True)
    masks = _sam_like_masks(M, H, W, 4).to(dev)

    def step():
        feat.grad = None; img.grad = None
        mean = mask_feature_mean(feat, masks, image_mask=img)
        loss = separation_loss(mean, 1000) + 0.1 * cohesion_loss(feat, masks, mean)
        loss.backward()
        return loss

    for _ in range(3):
        step()
    _lib.profile_enable(True); _lib.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 20
    for _ in range(n):
        step()
    e1.record(); torch.cuda.synchronize()
    pr = _lib.profile_read(); _lib.profile_enable(False)
    ms, cnt = pr["mask_stats"]
    traffic = 4 * (M * H * W + 7 * H * W * 4) + 2 * 6 * H * W * 4
    print(f"Stage-1 loss fwd+bwd, {M} masks {W}x{H} C={C}: {e0.elapsed_time(e1) / n:.3f} ms/step on the device "
          f"({ms / n:.3f} ms in the 4 streaming kernels, {traffic / (ms / n * 1e-3) / 1e9:.0f} GB/s of algorithmic bytes)")


if __name__ == "__main__":
    main()
