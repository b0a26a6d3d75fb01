"""graphs.GraphedViewStep: the Stage-1 training step of one camera (render + SAM-mask statistics + cohesion / separation
losses + backward, train.py:352-498) captured as a CUDA graph over the camera's resident geometry and replayed.  A
replay must give what the eager step gives on the same parameter values, see in-place parameter updates, and start over
when the geometry or the parameter tensors change."""
import types

import pytest
import torch

from opengaussian_b200 import synth

pytestmark = pytest.mark.gpu

PIPE = types.SimpleNamespace(debug=False, compute_cov3D_python=False, convert_SHs_python=False)


def _setup(scene, n_views, masks=12):
    from opengaussian_b200.mask_stats import cohesion_loss, get_SAM_mask_and_feat, mask_feature_mean, separation_loss
    from opengaussian_b200.renderer import render
    dev = torch.device("cuda")
    gs, cams = synth.make_scene(scene, n_views=n_views)
    pc = synth.SynthModel(gs, dev, stage0=False)
    cam = [types.SimpleNamespace(FoVx=c.FoVx, FoVy=c.FoVy, image_height=c.image_height, image_width=c.image_width,
                                 world_view_transform=c.world_view_transform.to(dev),
                                 full_proj_transform=c.full_proj_transform.to(dev),
                                 camera_center=c.camera_center.to(dev), bClusterOccur=None) for c in cams]
    H, W = cam[0].image_height, cam[0].image_width
    sam = [synth.sam_like_id_map(masks, H, W, 4 + v).to(dev) for v in range(n_views)]
    bg = torch.zeros(3, device=dev)

    def view_loss(i):
        out = render(cam[i], pc, PIPE, bg, 40_000, rescale=False)
        _, m, _ = get_SAM_mask_and_feat(sam[i], level=0, num_mask=masks)
        mean = mask_feature_mean(out["ins_feat"], m, image_mask=out["silhouette"])
        return separation_loss(mean, 40_000) + 0.1 * cohesion_loss(out["ins_feat"], m, mean)

    return dev, pc, view_loss


def _eager(view_loss, pc, i):
    pc._ins_feat.grad = None
    loss = view_loss(i)
    loss.backward()
    return float(loss), pc._ins_feat.grad.clone()


def _check(got_loss, got_grad, want):
    assert abs(float(got_loss) - want[0]) <= 1e-5 * abs(want[0]) + 1e-7
    assert float((got_grad - want[1]).abs().max()) <= 5e-5 * float(want[1].abs().max()) + 1e-12


@pytest.mark.parametrize("scene", ["plumbing_10k_256", "blender_300k_800"])
def test_graphed_step_matches_eager_and_tracks_parameters(scene):
    from opengaussian_b200 import rasterizer as rz
    from opengaussian_b200.graphs import GraphedViewStep, geometry_guard
    dev, pc, view_loss = _setup(scene, 2)
    rz.view_cache.clear()
    rz.view_cache.enabled = True
    step = GraphedViewStep(view_loss, [pc._ins_feat], guard=geometry_guard(pc))
    gen = torch.Generator(device=dev).manual_seed(3)
    want = [_eager(view_loss, pc, i) for i in (0, 1)]
    assert float(want[0][1].abs().max()) > 0
    for visit in range(3):                       # eager, capture + replay, replay
        for i in (0, 1):
            loss = step(i)
            _check(loss, pc._ins_feat.grad, want[i])
    assert step.stats() == dict(graphs=2, replays=4, captures=2, eager=2)

    # an optimizer step writes the parameter in place: the replay reads the new values
    for _ in range(2):
        with torch.no_grad():
            pc._ins_feat.add_(0.2 * torch.randn(pc._ins_feat.shape, device=dev, generator=gen))
        for i in (1, 0):
            w = _eager(view_loss, pc, i)
            loss = step(i)
            _check(loss, pc._ins_feat.grad, w)
    assert step.stats()["captures"] == 2 and step.stats()["replays"] == 8

    # the geometry moves (version counter): graphs are dropped, the next visits are eager / capture again
    with torch.no_grad():
        pc._opacity.sub_(0.5)
    w = _eager(view_loss, pc, 0)
    assert abs(w[0] - want[0][0]) > 0
    for visit in range(3):
        loss = step(0)
        _check(loss, pc._ins_feat.grad, w)
    assert step.stats()["captures"] == 3 and step.stats()["graphs"] == 1

    # a replaced parameter tensor: the old graphs point at the old storage and must not be replayed
    old = pc._ins_feat
    pc._ins_feat = (old.detach() * 0.5).requires_grad_(True)
    step.params = [pc._ins_feat]
    w = _eager(view_loss, pc, 0)
    loss = step(0)
    _check(loss, pc._ins_feat.grad, w)
    assert step.stats()["graphs"] == 0
    rz.view_cache.clear()


def test_uncapturable_views_run_eagerly():
    """Without a resident view (cache off) the forward has to read the duplicate count back: no capture, same results."""
    from opengaussian_b200 import rasterizer as rz
    from opengaussian_b200.graphs import GraphedViewStep
    dev, pc, view_loss = _setup("plumbing_10k_256", 1)
    rz.view_cache.clear()
    rz.view_cache.enabled = False
    try:
        step = GraphedViewStep(view_loss, [pc._ins_feat])
        want = _eager(view_loss, pc, 0)
        for _ in range(4):
            loss = step(0)
            _check(loss, pc._ins_feat.grad, want)
        assert step.stats()["graphs"] == 0 and step.stats()["replays"] == 0 and step.stats()["eager"] >= 4
    finally:
        rz.view_cache.enabled = True
    # and with the cache back on the same object captures after all (the failed view is retried once the signature changes)
    step2 = GraphedViewStep(view_loss, [pc._ins_feat])
    for _ in range(3):
        _check(step2(0), pc._ins_feat.grad, want)
    assert step2.stats()["graphs"] == 1
    rz.view_cache.clear()


def test_multi_view_step_with_accumulate():
    """step(view, accumulate=True) adds the view's gradient to the existing .grad -- eager, capturing and replaying visits."""
    from opengaussian_b200 import rasterizer as rz
    from opengaussian_b200.graphs import GraphedViewStep, geometry_guard
    dev, pc, view_loss = _setup("plumbing_10k_256", 2)
    rz.view_cache.clear()
    rz.view_cache.enabled = True
    w0, w1 = _eager(view_loss, pc, 0), _eager(view_loss, pc, 1)
    step = GraphedViewStep(view_loss, [pc._ins_feat], guard=geometry_guard(pc))
    for visit in range(4):
        step(0)
        step(1, accumulate=True)
        _check(w0[0] + w1[0], pc._ins_feat.grad, (w0[0] + w1[0], w0[1] + w1[1]))
    assert step.stats()["graphs"] == 2 and step.stats()["replays"] == 6
    rz.view_cache.clear()
