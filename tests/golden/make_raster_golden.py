"""Generates tests/golden/raster_pieces_golden.npz by running the REFERENCE's own Python code for the
parts of the rasterizer path that exist in the reference tree (the CUDA rasterizer itself is an
un-vendored dependency):

* SH basis, colour offset and clamp: utils/sh_utils.py::eval_sh + the `clamp_min(sh2rgb + 0.5, 0)`
  of gaussian_renderer/__init__.py:90-96 (the `convert_SHs_python` branch, which is documented to
  be equivalent to what the rasterizer does with `shs`);
* 3D covariance packing and quaternion convention: utils/general_utils.py::build_scaling_rotation /
  build_rotation / strip_symmetric as used by GaussianModel.get_covariance
  (scene/gaussian_model.py:63-67, the `compute_cov3D_python` branch);
* camera matrices: utils/graphics_utils.py::getWorld2View2 / getProjectionMatrix composed exactly as
  scene/cameras.py:71-78 does (transposed world-view, transposed projection, their product,
  camera centre).

Run in the build container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_raster_golden.py

Inputs are regenerated from numpy's MT19937 stream by `inputs()`; only the reference OUTPUTS are stored.
"""
import importlib.util
import math
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def _load(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def inputs():
    rs = np.random.RandomState(2024)
    P = 96
    xyz = ((rs.rand(P, 3) - 0.5) * 4.0).astype(np.float32)
    shs = (rs.randn(P, 16, 3) * 0.4).astype(np.float32)
    scales = np.exp(rs.randn(P, 3) * 0.5 - 2.0).astype(np.float32)
    rot = rs.randn(P, 4).astype(np.float32)
    rot /= np.linalg.norm(rot, axis=1, keepdims=True)
    campos = np.array([0.3, -5.0, 1.2], np.float32)
    # cameras: (R [3,3] camera-to-world as the dataset readers store it, T [3], FoVx, FoVy)
    cams = []
    for i in range(4):
        a = rs.randn(3, 3)
        q, _ = np.linalg.qr(a)
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        cams.append((q.astype(np.float64), (rs.randn(3) * 2).astype(np.float64), 0.69 + 0.1 * i, 0.55 + 0.07 * i))
    return dict(xyz=xyz, shs=shs, scales=scales, rot=rot, campos=campos, cams=cams)


def main():
    sh_utils = _load("utils/sh_utils.py", "ref_sh_utils")
    gfx = _load("utils/graphics_utils.py", "ref_graphics_utils")
    # build_rotation / strip_lowerdiag hard-code device="cuda": drop the device argument on this CPU box
    real_zeros = torch.zeros
    torch.zeros = lambda *a, **k: real_zeros(*a, **{kk: vv for kk, vv in k.items() if kk != "device"})
    try:
        gen = _load("utils/general_utils.py", "ref_general_utils")
        inp = inputs()
        out = {}
        xyz = torch.from_numpy(inp["xyz"])
        shs = torch.from_numpy(inp["shs"])
        campos = torch.from_numpy(inp["campos"])
        # --- SH -> RGB, reference lines gaussian_renderer/__init__.py:90-96 ---
        for deg in range(4):
            shs_view = shs.transpose(1, 2).view(-1, 3, 16)
            dir_pp = xyz - campos.repeat(xyz.shape[0], 1)
            dir_pp_normalized = dir_pp / dir_pp.norm(dim=1, keepdim=True)
            sh2rgb = sh_utils.eval_sh(deg, shs_view, dir_pp_normalized)
            out[f"rgb_deg{deg}"] = torch.clamp_min(sh2rgb + 0.5, 0.0).numpy()
            out[f"rgb_unclamped_deg{deg}"] = (sh2rgb + 0.5).numpy()
        # --- covariance, reference scene/gaussian_model.py:63-67 ---
        for mod in (1.0, 0.7):
            L = gen.build_scaling_rotation(mod * torch.from_numpy(inp["scales"]), torch.from_numpy(inp["rot"]))
            actual = L @ L.transpose(1, 2)
            out[f"cov3D_mod{mod}"] = gen.strip_symmetric(actual).numpy()
        # --- cameras, reference scene/cameras.py:71-78 ---
        for i, (R, T, fx, fy) in enumerate(inp["cams"]):
            wvt = torch.tensor(gfx.getWorld2View2(R, T, np.array([0.0, 0.0, 0.0]), 1.0)).transpose(0, 1)
            proj = gfx.getProjectionMatrix(znear=0.01, zfar=100.0, fovX=fx, fovY=fy).transpose(0, 1)
            full = (wvt.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0)
            center = wvt.inverse()[3, :3]
            out[f"cam{i}_world_view"] = wvt.numpy()
            out[f"cam{i}_proj"] = proj.numpy()
            out[f"cam{i}_full_proj"] = full.numpy()
            out[f"cam{i}_center"] = center.numpy()
            out[f"cam{i}_tan"] = np.array([math.tan(fx * 0.5), math.tan(fy * 0.5)], np.float64)
    finally:
        torch.zeros = real_zeros
    path = os.path.join(HERE, "raster_pieces_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.startswith(("rgb_deg3", "cov3D_mod1", "cam0"))})


if __name__ == "__main__":
    main()
