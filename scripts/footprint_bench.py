"""Developer timing probe: SAM-id votes of B single splats in one camera -- one batched call (csrc/footprint.cu)
vs the reference's per-splat flow (P = 1 rasterizer call + uint8 image + weighted bincount) on the same GPU.
Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opengaussian_b200 import synth  # noqa: E402
from opengaussian_b200.sam_footprints import batched_splat_ids, get_splat_id_and_weights  # noqa: E402


def main():
    dev = "cuda"
    gs, cams = synth.make_scene("scannet_1m_1296x968", n_views=2)
    cam = cams[0].to(dev)
    pc = synth.SynthModel(gs, dev)
    H, W = cam.image_height, cam.image_width
    rs = np.random.RandomState(0)
    blocks = rs.permutation(((H + 39) // 40) * ((W + 47) // 48)).reshape((H + 39) // 40, (W + 47) // 48)
    sam = torch.from_numpy(np.repeat(np.repeat(blocks, 40, axis=0), 48, axis=1)[:H, :W].astype(np.int64) - 1).to(dev)
    P = gs["means3D"].shape[0]
    res = {}
    for B in (1000, 100_000):
        ids = torch.arange(0, P, P // B, device=dev)[:B]
        batched_splat_ids(cam, pc, ids, sam)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 5
        for _ in range(n):
            out = batched_splat_ids(cam, pc, ids, sam)
        torch.cuda.synchronize()
        res[f"batched_ms_B{B}"] = round((time.perf_counter() - t0) / n * 1e3, 3)
        res[f"visible_B{B}"] = int(out["visible"].sum())
        res[f"mean_footprint_px_B{B}"] = round(float(out["footprint_pixels"].float().mean()), 1)
    ids = torch.arange(0, P, P // 1000, device=dev)[:50]
    get_splat_id_and_weights(cam, pc, int(ids[0]), sam)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for g in ids.tolist():
        get_splat_id_and_weights(cam, pc, g, sam)
    torch.cuda.synchronize()
    res["one_splat_flow_ms_per_splat"] = round((time.perf_counter() - t0) / len(ids) * 1e3, 3)
    print(json.dumps({"probe": "splat_footprint_votes", "image": [W, H], "gaussians": P, **res}))


if __name__ == "__main__":
    main()
