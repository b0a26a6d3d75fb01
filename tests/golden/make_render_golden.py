"""Generates tests/golden/render_golden.npz by running the REFERENCE's own `render()`
(/root/reference/gaussian_renderer/__init__.py:22-373, imported unchanged together with the reference's own
`scene/gaussian_model.py::GaussianModel`) on the CPU in the build container:

    python tests/golden/make_render_golden.py

What is stubbed, and why (the reference tree does not exist on the GPU box, so only OUTPUTS are committed):
  * `ashawkey_diff_gaussian_rasterization` -- the un-vendored CUDA rasterizer -- is replaced by a module whose
    `GaussianRasterizer` calls the CPU oracle (oracle/raster_oracle.c, forward and backward) once per pass.  Every
    pass the reference issues (RGB, ins_feat[:, :3], ins_feat[:, 3:6], silhouette, per-cluster, per-leaf) therefore
    runs through the reference's OWN control flow: filters, rescale draws, background handling, concatenations,
    thresholds, return dict.
  * `pytorch3d.ops.knn_points` (missing package; used only with post_process=True) -- brute-force torch version;
    `plyfile`, `bitarray` (missing packages, not used by render()) -- empty stand-ins so the imports succeed.
  * `.cuda()` / `device="cuda"` -- no-ops (CPU run).
The GPU test (tests/test_render_gpu.py::test_render_vs_reference_golden) feeds the same seeded inputs to
opengaussian_b200.renderer.render and compares every key of the returned dict.
"""
import math
import os
import sys
import types
from typing import NamedTuple

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
OUT = os.path.join(HERE, "render_golden.npz")

P, W, H = 3000, 96, 80
K1, K2 = 4, 3


def scene():
    """Seeded inputs (regenerated identically by the test): Gaussians, camera, cluster ids."""
    sys.path.insert(0, ROOT)
    from opengaussian_b200 import synth
    gs = synth.make_gaussians(P, "blender", 21, scale_mult=2.5)
    cam = synth.orbit_cameras(4, 3.5, W, H, 0.9, 0.8)[1]
    g = torch.Generator().manual_seed(77)
    cluster_idx = torch.randint(0, K1, (P,), generator=g)
    leaf_idx = cluster_idx * K2 + torch.randint(0, K2, (P,), generator=g)
    return gs, cam, cluster_idx, leaf_idx


def fill_model(pc, gs, with_q):
    """Sets the PARAMETER tensors of a GaussianModel-like object from activated synthetic values."""
    pc._xyz = gs["means3D"].clone()
    pc._scaling = torch.log(gs["scales"]).clone()
    pc._rotation = (gs["rotations"] * 1.3).clone()
    pc._opacity = torch.logit(gs["opacities"].clamp(1e-4, 1 - 1e-4)).clone()
    pc._features_dc = gs["shs"][:, :1].contiguous().clone()
    pc._features_rest = gs["shs"][:, 1:].contiguous().clone()
    pc._ins_feat = (gs["ins_feat"] * 2 - 1).clone().requires_grad_(True)
    pc._ins_feat_q = torch.empty(0)
    if with_q:
        g = torch.Generator().manual_seed(5)
        pc._ins_feat_q = torch.randn(gs["ins_feat"].shape, generator=g)
    pc.active_sh_degree = 3
    pc.max_sh_degree = 3
    return pc


# the cases: kwargs of render() (tensors are built in run_cases), the CPU RNG seed set right before the call
CASES = {
    "stage1": dict(seed=1, kw=dict(rescale=False)),
    "stage1_quantized_feat": dict(seed=2, with_q=True, kw=dict(rescale=False)),
    "stage2_rescaled": dict(seed=None, kw=dict(rescale=True)),                 # seed chosen so that prob > 0.5
    "stage2_not_rescaled": dict(seed=None, kw=dict(rescale=True)),             # seed chosen so that prob <= 0.5
    "stage22_cluster": dict(seed=3, kw=dict(rescale=False, render_feat_map=False, render_cluster=True,
                                            selected_root_id=2), clusters=True, leaves=True),
    "stage3_all_leaves": dict(seed=4, kw=dict(rescale=False, render_feat_map=False, render_color=True), leaves=True),
    "click_selected_leaf": dict(seed=5, kw=dict(rescale=False, render_feat_map=False, selected_root_id=1), leaves=True,
                                selected_leaf=[4, 5]),
    "better_vis_seg_rgb": dict(seed=6, kw=dict(rescale=False, render_feat_map=False, render_cluster=True, better_vis=True,
                                               seg_rgb=True, selected_root_id=0), clusters=True, leaves=True),
}


def seed_for(rescaled):
    """A CPU RNG seed whose first torch.rand(1) is > 0.5 (rescaled) or <= 0.5 (render(): :121-124)."""
    for s in range(100, 200):
        torch.manual_seed(s)
        if (float(torch.rand(1)) > 0.5) == rescaled:
            return s
    raise RuntimeError


def install_stubs():
    sys.path.insert(0, ROOT)
    from oracle import raster as orc

    class GaussianRasterizationSettings(NamedTuple):
        image_height: int
        image_width: int
        tanfovx: float
        tanfovy: float
        bg: torch.Tensor
        scale_modifier: float
        viewmatrix: torch.Tensor
        projmatrix: torch.Tensor
        sh_degree: int
        campos: torch.Tensor
        prefiltered: bool
        debug: bool

    flags_log = []

    class _Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, means3D, means2D, opacities, shs, colors_precomp, scales, rotations, cov3D, rs):
            n = lambda t: None if t is None else t.detach().numpy().astype(np.float32)  # noqa: E731
            cam = orc.Camera(W=rs.image_width, H=rs.image_height, tanfovx=rs.tanfovx, tanfovy=rs.tanfovy,
                             view=n(rs.viewmatrix).reshape(-1), proj=n(rs.projmatrix).reshape(-1), campos=n(rs.campos),
                             scale_modifier=float(rs.scale_modifier), sh_degree=int(rs.sh_degree))
            st = orc.forward(cam, n(means3D), n(opacities), n(scales), n(rotations), n(cov3D), n(shs), n(colors_precomp),
                             bg=n(rs.bg))
            ctx.st = st
            ctx.has = (shs is not None, colors_precomp is not None, scales is not None, cov3D is not None)
            flags_log.append(st.flags != 0)
            t = torch.from_numpy
            return (t(st.color.copy()), t(st.radii.astype(np.int32)), t(st.out_depth.copy())[None], t(st.out_alpha.copy())[None])

        @staticmethod
        def backward(ctx, g_color, _g_radii, g_depth, g_alpha):
            st = ctx.st
            z = lambda g, shape: np.zeros(shape, np.float32) if g is None else g.numpy().reshape(shape)  # noqa: E731
            Hh, Ww = st.cam.H, st.cam.W
            ref = orc.backward(st, z(g_color, (3, Hh, Ww)), z(g_depth, (Hh, Ww)), z(g_alpha, (Hh, Ww)))
            f = lambda a: None if a is None else torch.from_numpy(np.asarray(a, np.float32))  # noqa: E731
            has_sh, has_col, has_sc, has_cov = ctx.has
            return (f(ref["means3D"]), f(ref["means2D"]), f(ref["opacities"]), f(ref["shs"]) if has_sh else None,
                    f(ref["colors_precomp"]) if has_col else None, f(ref["scales"]) if has_sc else None,
                    f(ref["rotations"]) if has_sc else None, f(ref["cov3D_precomp"]) if has_cov else None, None)

    class GaussianRasterizer(torch.nn.Module):
        def __init__(self, raster_settings):
            super().__init__()
            self.raster_settings = raster_settings

        def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                    cov3D_precomp=None):
            return _Fn.apply(means3D, means2D, opacities, shs, colors_precomp, scales, rotations, cov3D_precomp,
                             self.raster_settings)

    mod = types.ModuleType("ashawkey_diff_gaussian_rasterization")
    mod.GaussianRasterizationSettings = GaussianRasterizationSettings
    mod.GaussianRasterizer = GaussianRasterizer
    sys.modules["ashawkey_diff_gaussian_rasterization"] = mod

    def knn_points(p1, p2, K=1, **kw):
        d = torch.cdist(p1[0], p2[0]) ** 2
        vals, idx = torch.topk(d, K, dim=1, largest=False)
        return types.SimpleNamespace(dists=vals[None], idx=idx[None])

    p3d = types.ModuleType("pytorch3d")
    p3d.ops = types.ModuleType("pytorch3d.ops")
    p3d.ops.knn_points = knn_points
    sys.modules["pytorch3d"], sys.modules["pytorch3d.ops"] = p3d, p3d.ops
    ply = types.ModuleType("plyfile")
    ply.PlyData = ply.PlyElement = object
    sys.modules["plyfile"] = ply
    ba = types.ModuleType("bitarray")
    ba.bitarray = object
    sys.modules["bitarray"] = ba
    # CPU run: `.cuda()` and device="cuda" are no-ops
    torch.Tensor.cuda = lambda self, *a, **k: self
    _zl, _tt = torch.zeros_like, torch.tensor

    def no_cuda(fn):
        def wrapped(*a, **k):
            if k.get("device") in ("cuda", torch.device("cuda")):
                k.pop("device")
            return fn(*a, **k)
        return wrapped

    torch.zeros_like = no_cuda(_zl)
    torch.tensor = no_cuda(_tt)
    return flags_log


def run_cases():
    flags_log = install_stubs()
    sys.path.insert(0, REF)
    import gaussian_renderer as ref_renderer                      # the reference's own render()
    from scene.gaussian_model import GaussianModel                # and its own parameter store / getters
    gs, cam, cluster_idx, leaf_idx = scene()
    pipe = types.SimpleNamespace(debug=False, compute_cov3D_python=False, convert_SHs_python=False)
    cam_ns = types.SimpleNamespace(FoVx=cam.FoVx, FoVy=cam.FoVy, image_height=cam.image_height, image_width=cam.image_width,
                                   world_view_transform=cam.world_view_transform, full_proj_transform=cam.full_proj_transform,
                                   camera_center=cam.camera_center, bClusterOccur=None)
    bg = torch.tensor([0.1, 0.3, 0.2])
    out = {}
    meta = {}
    for name, case in CASES.items():
        pc = fill_model(GaussianModel(3), gs, case.get("with_q", False))
        kw = dict(case["kw"])
        if case.get("clusters"):
            kw["cluster_idx"] = cluster_idx
        if case.get("leaves"):
            kw["leaf_cluster_idx"] = leaf_idx
        if case.get("selected_leaf") is not None:
            kw["selected_leaf_id"] = torch.tensor(case["selected_leaf"])
        seed = case["seed"]
        if seed is None:
            seed = seed_for(name == "stage2_rescaled")
        meta[name] = seed
        kw.update(root_num=K1, leaf_num=K2)
        del flags_log[:]
        torch.manual_seed(seed)
        res = ref_renderer.render(cam_ns, pc, pipe, bg, 1000, **kw)
        flagged = np.zeros((H, W), bool)
        for f in flags_log:
            flagged |= f
        out[f"{name}/flagged"] = flagged
        out[f"{name}/seed"] = np.int64(seed)
        for k, v in res.items():
            if v is None:
                out[f"{name}/{k}/none"] = np.int8(1)
            elif isinstance(v, torch.Tensor):
                out[f"{name}/{k}"] = v.detach().numpy()
            elif isinstance(v, list):
                out[f"{name}/{k}/len"] = np.int64(len(v))
                for i, t in enumerate(v):
                    out[f"{name}/{k}/{i}"] = t.detach().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
            else:
                out[f"{name}/{k}"] = np.asarray(v)
        # Stage-1 gradient: a loss on the feature map back to the PARAMETER _ins_feat through the reference's getters
        if name in ("stage1", "stage2_rescaled") and res["ins_feat"] is not None:
            g = torch.Generator().manual_seed(9)
            wgt = torch.randn(res["ins_feat"].shape, generator=g) * torch.from_numpy(~flagged).float()
            (res["ins_feat"] * wgt).sum().backward()
            out[f"{name}/grad_ins_feat"] = pc._ins_feat.grad.numpy().copy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v for k, v in meta.items()}, f"{os.path.getsize(OUT) / 1e6:.2f} MB")


if __name__ == "__main__":
    run_cases()
