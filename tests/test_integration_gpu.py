"""Integration on the GPU: the training-loop body of the reference (train.py:232-611) driven through the drop-ins --
`opengaussian_b200.renderer.render`, `mask_stats`, `Quantize_kMeans`, `FusedAdam` -- on a synthetic scene whose views
carry consistent blob masks.  The reference's drivers need a dataset on disk, `plyfile`, `pytorch3d` and a CUDA build
of the un-vendored rasterizer, none of which exist on the GPU box, so the loop body is restated here call for call
(the calls and their arguments are the reference's: render(...) at train.py:352-358, the Stage-1 losses at :448-456,
the Stage-2.1 codebook update at :322-332 and loss at :464-473, optimizer.step() at :608-610, render.py:61)."""
import types

import pytest
import torch

from opengaussian_b200 import synth

pytestmark = pytest.mark.gpu
PIPE = types.SimpleNamespace(debug=False, compute_cov3D_python=False, convert_SHs_python=False)


def _setup(P=60_000, W=320, H=240, views=4, blobs=12):
    dev = torch.device("cuda")
    gs = synth.make_gaussians(P, "blender", 3, scale_mult=0.8)
    cams = [c.to(dev) for c in synth.orbit_cameras(views, 4.0, W, H, 0.69, 1.0)]
    pc = synth.SynthModel(gs, dev, stage0=False)
    labels = synth.blob_labels(gs, blobs).to(dev)
    masks = [synth.blob_view_masks(c, pc, labels, blobs) for c in cams]
    assert all(m.shape[0] >= 3 for m in masks)
    return dev, gs, cams, pc, masks


def test_stage1_then_stage2_loop_trains():
    from opengaussian_b200.kmeans_quantize import Quantize_kMeans
    from opengaussian_b200.mask_stats import cohesion_loss, mask_feature_mean, separation_loss
    from opengaussian_b200.optim import FusedAdam
    from opengaussian_b200.renderer import render
    dev, gs, cams, pc, masks = _setup()
    bg = torch.zeros(3, device=dev)
    # gaussians.training_setup: Adam(l, lr=0.0, eps=1e-15) with the ins_feat group (scene/gaussian_model.py:215-230);
    # a larger step than the reference's 0.001 so that 120 iterations show what its 10 000 do
    opt = FusedAdam([{"params": [pc._ins_feat], "lr": 0.02, "name": "ins_feat"}], lr=0.0, eps=1e-15)
    feat0 = pc._ins_feat.detach().clone()
    losses = []
    for it in range(120):                                   # ---- Stage 1 (train.py:352-358, :448-456, :497, :608-610)
        v = it % len(cams)
        out = render(cams[v], pc, PIPE, bg, it, rescale=False)
        assert out["ins_feat"].shape == (6, cams[v].image_height, cams[v].image_width)
        mean = mask_feature_mean(out["ins_feat"], masks[v], image_mask=out["silhouette"])
        loss = separation_loss(mean, it) + 0.1 * cohesion_loss(out["ins_feat"], masks[v], mean)
        loss.backward()
        assert pc._ins_feat.grad is not None and pc._xyz.grad is None      # geometry is frozen (train.py:431-436)
        opt.step()
        opt.zero_grad(set_to_none=True)
        losses.append(loss.detach())
    losses = torch.stack(losses).cpu()
    first, last = float(losses[:12].mean()), float(losses[-12:].mean())
    print(f"stage-1 loss: first 12 steps {first:.5f} -> last 12 steps {last:.5f}")
    assert torch.isfinite(losses).all() and last < first
    assert float((pc._ins_feat.detach() - feat0).abs().max()) > 1e-3
    # ---- Stage 2.1: coarse codebook, quantised features rendered, masked L1 to the per-mask mean (train.py:322-332, :464-473)
    cb = Quantize_kMeans(num_clusters=16, num_leaf_clusters=4, num_iters=5, dim=9)
    for it in range(6):
        cb.forward(pc, it, assign=(it % 3 == 0), mode="root", pos_weight=0.5)
        assert pc._ins_feat_q.shape == pc._ins_feat.shape and cb.cls_ids.dtype == torch.int64
        v = it % len(cams)
        torch.manual_seed(it)
        out = render(cams[v], pc, PIPE, bg, it, rescale=True)               # quantised features, random rescale
        with torch.no_grad():
            mean = mask_feature_mean(out["ins_feat"], masks[v], image_mask=out["silhouette"])
            target = (masks[v].float()[:, None] * mean[:, :, None, None]).sum(0)
        covered = masks[v].any(0)
        loss = ((out["ins_feat"] - target).abs() * covered).sum() / covered.sum().clamp(min=1)
        loss.backward()
        g = pc._ins_feat.grad
        assert g is not None and bool(torch.isfinite(g).all()) and float(g.abs().max()) > 0   # straight-through (:273-275)
        opt.step()
        opt.zero_grad(set_to_none=True)
    assert int(cb.cls_ids.min()) >= 0 and int(cb.cls_ids.max()) < 16
    # ---- render.py:61: offline render of every view under no_grad
    with torch.no_grad():
        for c in cams:
            out = render(c, pc, PIPE, bg, 0, rescale=False)
            for k in ("render", "ins_feat", "alpha", "depth", "silhouette"):
                assert bool(torch.isfinite(out[k]).all())
            assert out["ins_feat"][:3].shape == out["render"].shape


def test_two_threads_two_streams():
    """The allocation callback state travels through the C ABI's alloc_user and the per-device statics are keyed by
    device: two Python threads rendering on two streams at the same time get the single-threaded results."""
    import threading
    from opengaussian_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer
    dev = torch.device("cuda")
    gs = synth.make_gaussians(20_000, "blender", 5, scale_mult=1.0)
    cams = [c.to(dev) for c in synth.orbit_cameras(2, 4.0, 256, 192, 0.69, 1.0)]
    t = {k: gs[k].to(dev) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
    bg = torch.zeros(3, device=dev)

    def frame(cam, params):
        rs = GaussianRasterizationSettings(cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy, bg, 1.0,
                                           cam.world_view_transform, cam.full_proj_transform, 3, cam.camera_center, False, False)
        out = GaussianRasterizer(rs)(means2D=torch.zeros_like(params["means3D"]), **params)
        (out[0].sum() + out[2].sum()).backward()
        return out[0].detach().clone(), params["shs"].grad.detach().clone()

    def leaves():
        return {k: v.clone().requires_grad_(True) for k, v in t.items()}

    want = [frame(c, leaves()) for c in cams]
    torch.cuda.synchronize()
    got, errs = [None, None], []

    def worker(i):
        try:
            s = torch.cuda.Stream(dev)
            with torch.cuda.stream(s):
                for _ in range(20):
                    got[i] = frame(cams[i], leaves())
            s.synchronize()
        except Exception as e:      # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errs, errs
    for i in range(2):
        assert torch.equal(got[i][0], want[i][0])
        assert float((got[i][1] - want[i][1]).abs().max()) <= 1e-5 * float(want[i][1].abs().max())
