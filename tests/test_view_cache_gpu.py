"""Frozen-geometry view reuse (opengaussian_b200.rasterizer.ViewCache, C ABI ogs_raster_forward_cached): from Stage 1 on
OpenGaussian trains `_ins_feat` only (train.py:431-436), so a camera's projected records and tile lists are computed once
and later render() calls of that camera run the blend kernel alone.  Everything a cached call returns must be
BIT-identical to what a fresh forward returns on the same inputs, and every way the geometry can legitimately change
(in-place update, new tensors, optimizer step) must miss."""
import types

import pytest
import torch

from opengaussian_b200 import synth

pytestmark = pytest.mark.gpu

PIPE = types.SimpleNamespace(debug=False, compute_cov3D_python=False, convert_SHs_python=False)
IMG_KEYS = ("render", "alpha", "depth", "silhouette", "ins_feat")


def _cam(c, dev):
    return types.SimpleNamespace(FoVx=c.FoVx, FoVy=c.FoVy, image_height=c.image_height, image_width=c.image_width,
                                 world_view_transform=c.world_view_transform.to(dev),
                                 full_proj_transform=c.full_proj_transform.to(dev),
                                 camera_center=c.camera_center.to(dev), bClusterOccur=None)


def _render(cam, pc, bg, G, cached):
    """One Stage-1 style render + backward; returns the dict's images, radii and dL/d_ins_feat."""
    from opengaussian_b200 import rasterizer as rz
    from opengaussian_b200.renderer import render
    pc._ins_feat.grad = None
    prev = rz.view_cache.enabled
    rz.view_cache.enabled = cached
    try:
        out = render(cam, pc, PIPE, bg, 40_000, rescale=False)
        (out["ins_feat"] * G).sum().backward()
    finally:
        rz.view_cache.enabled = prev
    return {k: out[k].detach().clone() for k in IMG_KEYS}, out["radii"].clone(), pc._ins_feat.grad.clone()


def _same(a, b):
    for k in IMG_KEYS:
        assert torch.equal(a[0][k], b[0][k]), k
    assert torch.equal(a[1], b[1])
    # the backward's per-(tile, Gaussian) reds land in a different order from run to run: rounding only
    assert float((a[2] - b[2]).abs().max()) <= 2e-5 * float(b[2].abs().max()) + 1e-12


@pytest.mark.parametrize("scene", ["plumbing_10k_256", "blender_300k_800"])
def test_cached_view_is_bit_identical_and_invalidates(scene):
    from opengaussian_b200 import rasterizer as rz
    dev = torch.device("cuda")
    gs, cams = synth.make_scene(scene, n_views=2)
    cam0, cam1 = _cam(cams[0], dev), _cam(cams[1], dev)
    pc = synth.SynthModel(gs, dev, stage0=False)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    gen = torch.Generator(device=dev).manual_seed(1)
    G = torch.randn(6, cam0.image_height, cam0.image_width, device=dev, generator=gen)
    vc = rz.view_cache
    vc.clear()
    h0, m0 = vc.hits, vc.misses

    ref0 = _render(cam0, pc, bg, G, cached=False)
    assert len(vc) == 0 and (vc.hits, vc.misses) == (h0, m0)          # disabled: untouched
    first = _render(cam0, pc, bg, G, cached=True)                     # miss: fills the entry
    assert len(vc) == 1 and vc.misses == m0 + 1
    _same(first, ref0)
    second = _render(cam0, pc, bg, G, cached=True)                    # hit: blend only
    assert vc.hits == h0 + 1
    _same(second, ref0)

    # the trained tensor changes every step: still a hit, and the new features are what gets composited
    with torch.no_grad():
        pc._ins_feat.add_(0.3 * torch.randn(pc._ins_feat.shape, device=dev, generator=gen))
    hit = _render(cam0, pc, bg, G, cached=True)
    assert vc.hits == h0 + 2
    ref1 = _render(cam0, pc, bg, G, cached=False)
    _same(hit, ref1)
    assert not torch.equal(hit[0]["ins_feat"], ref0[0]["ins_feat"])

    # train.py:431-436 re-detaches the geometry every iteration: new tensor objects, same storage and version -> hit
    pc._xyz, pc._scaling, pc._opacity = pc._xyz.detach(), pc._scaling.detach(), pc._opacity.detach()
    _same(_render(cam0, pc, bg, G, cached=True), ref1)
    assert vc.hits == h0 + 3

    # a second camera gets its own entry
    ref_c1 = _render(cam1, pc, bg, G, cached=False)
    _same(_render(cam1, pc, bg, G, cached=True), ref_c1)
    _same(_render(cam1, pc, bg, G, cached=True), ref_c1)
    assert len(vc) == 2 and vc.hits == h0 + 4

    # an in-place edit of the geometry bumps the version counter: miss, the entry is replaced, results follow
    with torch.no_grad():
        pc._xyz.add_(0.01)
    m1 = vc.misses
    moved = _render(cam0, pc, bg, G, cached=True)
    assert vc.misses == m1 + 1 and len(vc) == 2
    ref2 = _render(cam0, pc, bg, G, cached=False)
    _same(moved, ref2)
    assert not torch.equal(ref2[0]["alpha"], ref1[0]["alpha"])
    _same(_render(cam0, pc, bg, G, cached=True), ref2)                # and hits again afterwards

    # new parameter tensors (densification, load_ply): miss
    pc._opacity = pc._opacity.clone()
    m2 = vc.misses
    _same(_render(cam0, pc, bg, G, cached=True), ref2)
    assert vc.misses == m2 + 1
    vc.clear()
    assert len(vc) == 0 and vc.bytes == 0


def test_cache_budget_evicts_least_recently_used():
    from opengaussian_b200 import rasterizer as rz
    dev = torch.device("cuda")
    gs, cams = synth.make_scene("plumbing_10k_256", n_views=3)
    cam = [_cam(c, dev) for c in cams]
    pc = synth.SynthModel(gs, dev, stage0=False)
    bg = torch.zeros(3, device=dev)
    G = torch.ones(6, cam[0].image_height, cam[0].image_width, device=dev)
    vc = rz.view_cache
    vc.clear()
    old_budget = vc.max_bytes
    try:
        ref = [_render(c, pc, bg, G, cached=False) for c in cam]
        _render(cam[0], pc, bg, G, cached=True)
        one = vc.bytes
        assert one > 0
        vc.max_bytes = int(2.5 * one)                                  # room for two of the three views
        ev0 = vc.evictions
        for i in (1, 2, 0, 1, 2):
            _same(_render(cam[i], pc, bg, G, cached=True), ref[i])
        assert len(vc) == 2 and vc.bytes <= vc.max_bytes and vc.evictions > ev0
        vc.max_bytes = 1                                               # nothing fits: plain forwards, still right
        vc.clear()
        _same(_render(cam[0], pc, bg, G, cached=True), ref[0])
        assert len(vc) == 0
    finally:
        vc.max_bytes = old_budget
        vc.clear()


def test_trainable_geometry_is_never_cached():
    from opengaussian_b200 import rasterizer as rz
    from opengaussian_b200.renderer import render
    dev = torch.device("cuda")
    gs, cams = synth.make_scene("plumbing_10k_256", n_views=1)
    cam = _cam(cams[0], dev)
    pc = synth.SynthModel(gs, dev, stage0=True)
    rz.view_cache.clear()
    for _ in range(2):
        out = render(cam, pc, PIPE, torch.zeros(3, device=dev), 100, rescale=False)
        out["render"].sum().backward()
    assert len(rz.view_cache) == 0
    with torch.no_grad():                                              # evaluation renders of a trainable model: not cached either
        render(cam, pc, PIPE, torch.zeros(3, device=dev), 100, rescale=False)
    assert len(rz.view_cache) == 0


def test_reference_call_convention_on_a_cached_view():
    """The reference's own render() (gaussian_renderer/__init__.py:104-163) hands the rasterizer ACTIVATED tensors,
    `colors_precomp = ins_feat[:, :3]` and a means2D that requires grad: the cached state must serve the full geometry
    backward (dL/dmeans2D) and a colour input that changes between calls."""
    from opengaussian_b200 import rasterizer as rz
    from opengaussian_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer
    import math
    dev = torch.device("cuda")
    gs, cams = synth.make_scene("plumbing_10k_256", n_views=1)
    c = _cam(cams[0], dev)
    rs = GaussianRasterizationSettings(c.image_height, c.image_width, math.tan(c.FoVx * 0.5), math.tan(c.FoVy * 0.5),
                                       torch.tensor([0.3, 0.2, 0.1], device=dev), 1.0, c.world_view_transform,
                                       c.full_proj_transform, 3, c.camera_center, False, False)
    geo = {k: gs[k].to(dev) for k in ("means3D", "opacities", "scales", "rotations")}
    shs = gs["shs"].to(dev)
    gen = torch.Generator(device=dev).manual_seed(2)
    G = torch.randn(3, c.image_height, c.image_width, device=dev, generator=gen)

    def run(colors, cached):
        prev, rz.view_cache.enabled = rz.view_cache.enabled, cached
        try:
            m2 = torch.zeros_like(geo["means3D"], requires_grad=True)
            col = colors.clone().requires_grad_(True)
            img, radii, depth, alpha = GaussianRasterizer(rs)(means2D=m2, colors_precomp=col, **geo)
            ((img * G).sum() + depth.sum() + 2 * alpha.sum()).backward()
            return img.detach(), radii, depth.detach(), alpha.detach(), m2.grad, col.grad
        finally:
            rz.view_cache.enabled = prev

    rz.view_cache.clear()
    f0 = gs["ins_feat"][:, :3].to(dev)
    f1 = gs["ins_feat"][:, 3:].to(dev)
    ref0, ref1 = run(f0, False), run(f1, False)
    h0 = rz.view_cache.hits
    z = run(f0, True)                                                   # activated tensors: the first sighting is not kept
    assert len(rz.view_cache) == 0
    a, b, cc = run(f0, True), run(f0, True), run(f1, True)              # admitted (same tensors again), hit, hit with other colours
    assert len(rz.view_cache) == 1 and rz.view_cache.hits == h0 + 2
    for got, want in ((z, ref0), (a, ref0), (b, ref0), (cc, ref1)):
        for x, y in zip(got[:4], want[:4]):
            assert torch.equal(x, y)
        for x, y in zip(got[4:], want[4:]):
            assert float((x - y).abs().max()) <= 2e-5 * float(y.abs().max()) + 1e-12
    # an SH pass of the same camera and geometry is its own entry (SH colours live in the cached records)
    img_sh = GaussianRasterizer(rs)(means2D=torch.zeros_like(geo["means3D"]), shs=shs, **geo)[0]
    img_sh2 = GaussianRasterizer(rs)(means2D=torch.zeros_like(geo["means3D"]), shs=shs, **geo)[0]
    img_sh3 = GaussianRasterizer(rs)(means2D=torch.zeros_like(geo["means3D"]), shs=shs, **geo)[0]
    assert len(rz.view_cache) == 2 and torch.equal(img_sh3, img_sh)
    # per-call temporaries (what the reference's getters produce) never become entries
    for _ in range(3):
        tmp = {k: v.clone() for k, v in geo.items()}
        GaussianRasterizer(rs)(means2D=torch.zeros_like(geo["means3D"]), shs=shs, **tmp)
    assert len(rz.view_cache) == 2
    rz.view_cache.enabled = False
    try:
        img_sh_ref = GaussianRasterizer(rs)(means2D=torch.zeros_like(geo["means3D"]), shs=shs, **geo)[0]
    finally:
        rz.view_cache.enabled = True
    assert torch.equal(img_sh, img_sh_ref) and torch.equal(img_sh2, img_sh_ref)
    assert torch.equal(run(f0, True)[0], ref0[0])                       # the precomputed-colour entry is still intact
    rz.view_cache.clear()


@pytest.mark.parametrize("cached", [False, True])
def test_multi_view_step_accumulates_feature_gradients(cached):
    """A Stage-1 step over several views (dist.render_views_backward): the 2nd..Vth view add their `_ins_feat` gradient
    into the existing .grad inside the backward's own kernel (feat_grad_kernel, accumulate form) -- with resident views
    and without.  Must equal the sum of the views' separate gradients."""
    from opengaussian_b200 import dist as ogd, rasterizer as rz
    from opengaussian_b200.renderer import render
    dev = torch.device("cuda")
    gs, cams = synth.make_scene("plumbing_10k_256", n_views=3)
    cam = [_cam(c, dev) for c in cams]
    pc = synth.SynthModel(gs, dev, stage0=False)
    bg = torch.zeros(3, device=dev)
    gen = torch.Generator(device=dev).manual_seed(9)
    G = [torch.randn(6, cam[0].image_height, cam[0].image_width, device=dev, generator=gen) for _ in cam]
    rz.view_cache.clear()
    want = torch.zeros_like(pc._ins_feat)
    for i in range(3):
        want += _render(cam[i], pc, bg, G[i], cached=False)[2]
    prev, rz.view_cache.enabled = rz.view_cache.enabled, cached
    try:
        for rep in range(2):                 # second repetition: every view is resident when cached
            pc._ins_feat.grad = None
            ogd.render_views_backward(lambda i: (render(cam[i], pc, PIPE, bg, 40_000, rescale=False)["ins_feat"] * G[i]).sum(),
                                      [0, 1, 2], [pc._ins_feat], already_split=True)
            got = pc._ins_feat.grad
            assert float((got - want).abs().max()) <= 3e-5 * float(want.abs().max())
        assert len(rz.view_cache) == (3 if cached else 0)
    finally:
        rz.view_cache.enabled = prev
        rz.view_cache.clear()


def test_pipe_flag_keeps_a_call_out_of_the_cache():
    from opengaussian_b200 import rasterizer as rz
    from opengaussian_b200.renderer import render
    dev = torch.device("cuda")
    gs, cams = synth.make_scene("plumbing_10k_256", n_views=1)
    cam = _cam(cams[0], dev)
    pc = synth.SynthModel(gs, dev, stage0=False)
    bg = torch.zeros(3, device=dev)
    rz.view_cache.clear()
    off = types.SimpleNamespace(debug=False, compute_cov3D_python=False, convert_SHs_python=False, view_cache=False)
    a = render(cam, pc, off, bg, 40_000, rescale=False)
    assert len(rz.view_cache) == 0 and rz.view_cache.enabled
    b = render(cam, pc, PIPE, bg, 40_000, rescale=False)
    c = render(cam, pc, PIPE, bg, 40_000, rescale=False)
    assert len(rz.view_cache) == 1
    for k in IMG_KEYS:
        assert torch.equal(a[k], b[k]) and torch.equal(a[k], c[k])
    b["radii"].add_(1)                                  # the first call's radii belong to the caller ...
    d = render(cam, pc, PIPE, bg, 40_000, rescale=False)
    assert torch.equal(d["radii"], a["radii"])          # ... the entry keeps its own copy
    rz.view_cache.clear()
