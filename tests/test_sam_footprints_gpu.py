"""Batched single-splat footprint votes (csrc/footprint.cu through the C ABI) against
* the one-splat flow of the reference (render_single_gaussian -> fix_image -> rgb_to_weight_map -> weighted
  bincount, utils/sam_refinement_utils.py:902-913) run on the B200 rasterizer: the uint8 images must agree pixel
  for pixel, so footprint size, uint8 maximum and the integer vote are compared EXACTLY;
* the golden values the reference's own post-processing produced from the CPU oracle's renders: the oracle uses
  expf where the kernels use ex2.approx, so single uint8 steps may differ -- footprint sizes within 1 % + 2 pixels,
  dominant ids equal except on near-ties (at most 2 % of the splats)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from opengaussian_b200 import synth

pytestmark = pytest.mark.gpu

FGOLD = os.path.join(os.path.dirname(__file__), "golden", "footprint_golden.npz")


def _gm():
    spec = importlib.util.spec_from_file_location("mfg", os.path.join(os.path.dirname(FGOLD), "make_footprint_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _one_splat_reference(cam, pc, gid, sam):
    """The reference flow for one splat, plus the integer statistics of its uint8 image."""
    from opengaussian_b200 import sam_footprints as sf
    img, radii, _, _ = sf.render_single_gaussian(cam, pc, gid, use_view_inv_white_shs=True)
    q = sf.fix_image(img.detach())
    assert torch.equal(q[..., 0], q[..., 1]) and torch.equal(q[..., 0], q[..., 2])
    q = q[..., 0].long()
    lo = int(sam.min())
    counts = torch.bincount((sam.long() - lo).flatten(), weights=q.flatten().double())
    k = int(torch.argmax(counts))
    hi = int(sam.max())
    empty_id = lo if (lo < 0 or lo == hi) else 0        # what the reference's argmax over all-zero counts returns
    did_float, _, vis = sf.get_splat_id_and_weights(cam, pc, gid, sam)
    return dict(pixels=int((q > 0).sum()), q_max=int(q.max()), dominant=k + lo if int(q.max()) > 0 else empty_id,
                weight=int(counts[k]), radius=int(radii[0]), float_flow_id=did_float, visible=vis, counts=counts)


@pytest.mark.parametrize("W,H", [(256, 256), (203, 131)])
def test_batched_votes_equal_the_one_splat_flow(W, H):
    from opengaussian_b200.sam_footprints import batched_splat_ids
    dev = "cuda"
    gs = synth.make_gaussians(4000, "blender", 1, scale_mult=0.6)
    cam = synth.orbit_cameras(3, 4.0, W, H, 0.69, 1.0)[1].to(dev)
    pc = synth.SynthModel(gs, dev)
    rs = np.random.RandomState(7)
    blocks = rs.permutation(((H + 12) // 13) * ((W + 16) // 17)).reshape((H + 12) // 13, (W + 16) // 17) % 37
    sam = torch.from_numpy(np.repeat(np.repeat(blocks, 13, axis=0), 17, axis=1)[:H, :W].astype(np.int64) - 1).to(dev)
    ids = torch.arange(5, 4000, 53, device=dev)                       # 76 splats, some behind / outside the view
    out = batched_splat_ids(cam, pc, ids, sam)
    n_vis = 0
    for j, gid in enumerate(ids.tolist()):
        ref = _one_splat_reference(cam, pc, gid, sam)
        assert int(out["radii"][j]) == ref["radius"], gid
        assert int(out["footprint_pixels"][j]) == ref["pixels"], gid
        assert int(out["q_max"][j]) == ref["q_max"], gid
        assert bool(out["visible"][j]) == ref["visible"], gid
        assert int(out["dominant_id"][j]) == ref["dominant"], gid
        assert int(out["dominant_weight"][j]) == (ref["weight"] if ref["pixels"] else 0), gid
        if ref["float_flow_id"] != ref["dominant"]:                   # float32 atomics vs exact integers: only on near-ties
            top2 = torch.topk(ref["counts"], 2).values
            assert float(top2[0] - top2[1]) <= 1e-4 * float(top2[0]), gid
        n_vis += ref["visible"]
    assert n_vis >= 20


def test_unique_ids_per_pixel_and_the_overflow_fallback():
    """Every pixel its own id: the dominant id is the FIRST pixel holding the uint8 maximum (argmax order), and any
    footprint above 128 pixels overflows the per-warp table and is redone by the one-splat path."""
    from opengaussian_b200.sam_footprints import batched_splat_ids, get_splat_id_and_weights
    dev = "cuda"
    W, H = 128, 96
    gs = synth.make_gaussians(600, "blender", 2, scale_mult=0.25)
    cam = synth.orbit_cameras(2, 4.0, W, H, 0.69, 1.0)[0].to(dev)
    pc = synth.SynthModel(gs, dev)
    sam = torch.arange(H * W, device=dev).view(H, W) + 5
    ids = torch.arange(0, 600, 9, device=dev)
    out = batched_splat_ids(cam, pc, ids, sam)
    small = large = 0
    for j, gid in enumerate(ids.tolist()):
        did, _, vis = get_splat_id_and_weights(cam, pc, gid, sam)
        assert bool(out["visible"][j]) == vis
        assert int(out["dominant_id"][j]) == did, (gid, int(out["footprint_pixels"][j]))
        n = int(out["footprint_pixels"][j])
        small += 0 < n <= 128
        large += n > 128
    assert large > 0 and small + large > 10


@pytest.mark.parametrize("name", ["ids_from_minus1", "ids_from_3"])
def test_batched_votes_vs_reference_golden(name):
    from opengaussian_b200.sam_footprints import batched_splat_ids
    dev = "cuda"
    m, gold = _gm(), np.load(FGOLD)
    gs, cam, sam = m.inputs(name)
    pc = synth.SynthModel(gs, dev)
    P = gs["means3D"].shape[0]
    out = batched_splat_ids(cam.to(dev), pc, torch.arange(P, device=dev), torch.from_numpy(sam).to(dev))
    px, gpx = out["footprint_pixels"].cpu().numpy().astype(np.int64), gold[f"{name}/footprint_pixels"]
    assert np.all(np.abs(px - gpx) <= 0.01 * gpx + 2), np.abs(px - gpx).max()
    assert np.all(np.abs(out["q_max"].cpu().numpy() - gold[f"{name}/q_max"]) <= 1)
    assert np.array_equal(out["visible"].cpu().numpy(), gold[f"{name}/visible"])
    mism = (out["dominant_id"].cpu().numpy() != gold[f"{name}/dominant_id"]).mean()
    assert mism <= 0.02, mism


def test_footprint_edge_cases():
    from opengaussian_b200 import _lib
    from opengaussian_b200.sam_footprints import batched_splat_ids
    dev = "cuda"
    gs = synth.make_gaussians(50, "blender", 3, scale_mult=0.5)
    cam = synth.orbit_cameras(2, 4.0, 64, 48, 0.69, 1.0)[0].to(dev)
    pc = synth.SynthModel(gs, dev)
    sam = torch.full((48, 64), 7, device=dev)
    empty = batched_splat_ids(cam, pc, torch.zeros(0, dtype=torch.long, device=dev), sam)
    assert empty["dominant_id"].shape == (0,) and empty["visible"].shape == (0,)
    out = batched_splat_ids(cam, pc, torch.arange(50, device=dev), sam)
    assert bool((out["dominant_id"] == 7).all())                       # one id everywhere (reference :680-681)
    pc._xyz.data[:, :] = cam.camera_center + 100.0 * (cam.camera_center / cam.camera_center.norm())   # behind the camera
    out = batched_splat_ids(cam, pc, torch.arange(50, device=dev), sam - 9)
    assert not bool(out["visible"].any()) and bool((out["dominant_id"] == -2).all()) and bool((out["radii"] == 0).all())
    with pytest.raises(_lib.OgsError):
        batched_splat_ids(cam, pc, torch.arange(3), sam.cpu())
