"""Shared helpers for the parity tests: small seeded scenes in oracle (numpy) form."""
import math

import numpy as np
import torch

from opengaussian_b200 import synth
from oracle import raster as orc


def small_scene(P=300, W=64, H=48, seed=0, kind="blender", fovx=0.9, radius=3.5, scale_mult=2.5, view=0,
                n_views=4, height=0.8):
    gs = synth.make_gaussians(P, kind, seed, scale_mult=scale_mult)
    cams = synth.orbit_cameras(n_views, radius, W, H, fovx, height)
    return gs, cams[view]


def to_oracle_cam(cam, sh_degree=3, scale_modifier=1.0):
    return orc.Camera(W=cam.image_width, H=cam.image_height, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
                      view=cam.world_view_transform.numpy().reshape(-1).copy(),
                      proj=cam.full_proj_transform.numpy().reshape(-1).copy(),
                      campos=cam.camera_center.numpy().copy(), scale_modifier=scale_modifier,
                      sh_degree=sh_degree)


def np_inputs(gs):
    return {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in gs.items()}


def grad_violations(got, want, rtol=1e-3):
    """Per-element gradient check (BASELINE north_star: "gradients within 1e-3 relative"): an element passes when
    |got - want| <= rtol * |want| + rtol * rms(want).  Returns (violating fraction, worst |diff| / tolerance)."""
    got = np.asarray(got, np.float64).reshape(-1)
    want = np.asarray(want, np.float64).reshape(-1)
    if want.size == 0:
        return 0.0, 0.0
    rms = float(np.sqrt(np.mean(want * want)))
    tol = rtol * np.abs(want) + rtol * rms + 1e-30
    ratio = np.abs(got - want) / tol
    return float((ratio > 1.0).mean()), float(ratio.max())
