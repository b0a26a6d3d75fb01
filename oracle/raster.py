"""ctypes front-end of oracle/raster_oracle.c (TEST INFRASTRUCTURE ONLY).

``forward`` / ``backward`` mirror one ``GaussianRasterizer`` call of the reference
(gaussian_renderer/__init__.py:104-112) on numpy arrays, with an optional block of extra
per-Gaussian feature channels composited beside RGB (OpenGaussian's ``ins_feat``).
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import lib

_f = np.float32


def _p(a, t=C.c_void_p):
    return None if a is None else a.ctypes.data_as(t)


def _c(a, dt=_f):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


@dataclass
class Camera:
    W: int
    H: int
    tanfovx: float
    tanfovy: float
    view: np.ndarray   # raw 16 floats of the transposed world->view tensor
    proj: np.ndarray   # raw 16 floats of the transposed full projection tensor
    campos: np.ndarray
    scale_modifier: float = 1.0
    sh_degree: int = 3


@dataclass
class ForwardState:
    cam: Camera
    P: int
    C: int
    radii: np.ndarray
    xy: np.ndarray
    depth: np.ndarray
    cov3D: np.ndarray
    conic_opacity: np.ndarray
    rgb: np.ndarray
    clamped: np.ndarray
    tiles_touched: np.ndarray
    offsets: np.ndarray
    N: int
    keys: np.ndarray
    point_list: np.ndarray
    ranges: np.ndarray
    colors: np.ndarray          # [P, C] as blended
    bg: np.ndarray
    color: np.ndarray           # [C, H, W]
    out_depth: np.ndarray
    out_alpha: np.ndarray
    final_T: np.ndarray
    n_contrib: np.ndarray
    flags: np.ndarray
    inputs: dict = field(default_factory=dict)


def preprocess(cam, means3D, opacities, scales=None, rotations=None, cov3D_precomp=None, shs=None):
    L = lib()
    P = means3D.shape[0]
    means3D = _c(means3D); opacities = _c(opacities).reshape(-1)
    scales = _c(scales); rotations = _c(rotations); cov3D_precomp = _c(cov3D_precomp); shs = _c(shs)
    M = 0 if shs is None else shs.shape[1]
    radii = np.zeros(P, np.int32); xy = np.zeros((P, 2), _f); depth = np.zeros(P, _f)
    cov3D = np.zeros((P, 6), _f); co = np.zeros((P, 4), _f)
    rgb = np.zeros((P, 3), _f); clamped = np.zeros((P, 3), np.uint8); tiles = np.zeros(P, np.uint32)
    L.ogs_oracle_preprocess(
        C.c_int(P), C.c_int(cam.sh_degree), C.c_int(M), _p(means3D), _p(scales), _p(rotations),
        _p(cov3D_precomp), _p(opacities), _p(shs), C.c_float(cam.scale_modifier),
        _p(_c(cam.view).reshape(-1)), _p(_c(cam.proj).reshape(-1)), _p(_c(cam.campos)),
        C.c_int(cam.W), C.c_int(cam.H), C.c_float(cam.tanfovx), C.c_float(cam.tanfovy),
        _p(radii), _p(xy), _p(depth), _p(cov3D), _p(co),
        _p(rgb) if shs is not None else None, _p(clamped) if shs is not None else None, _p(tiles))
    return radii, xy, depth, cov3D, co, rgb, clamped, tiles


def bin_tiles(cam, radii, xy, depth, tiles):
    L = lib()
    P = radii.shape[0]
    offsets = np.zeros(P, np.uint32)
    L.ogs_oracle_scan.restype = C.c_int64
    N = int(L.ogs_oracle_scan(C.c_int(P), _p(tiles), _p(offsets)))
    gx, gy = (cam.W + 15) // 16, (cam.H + 15) // 16
    keys = np.zeros(max(N, 1), np.uint64); vals = np.zeros(max(N, 1), np.uint32)
    ranges = np.zeros((gx * gy, 2), np.uint32)
    L.ogs_oracle_bin(C.c_int(P), C.c_int(cam.W), C.c_int(cam.H), _p(xy), _p(depth), _p(radii),
                     _p(offsets), C.c_int64(N), _p(keys), _p(vals), _p(ranges))
    return offsets, N, keys[:N], vals[:N], ranges


def forward(cam, means3D, opacities, scales=None, rotations=None, cov3D_precomp=None, shs=None,
            colors_precomp=None, extra=None, bg=None, margin=1e-4) -> ForwardState:
    """extra: optional [P, F] feature channels appended after the 3 colour channels."""
    L = lib()
    P = means3D.shape[0]
    assert (shs is None) != (colors_precomp is None)
    radii, xy, depth, cov3D, co, rgb, clamped, tiles = preprocess(
        cam, means3D, opacities, scales, rotations, cov3D_precomp, shs)
    offsets, N, keys, plist, ranges = bin_tiles(cam, radii, xy, depth, tiles)
    base = rgb if shs is not None else _c(colors_precomp)
    colors = base if extra is None else np.concatenate([base, _c(extra)], axis=1)
    colors = _c(colors)
    Cn = colors.shape[1]
    bg = np.zeros(Cn, _f) if bg is None else _c(bg)
    if bg.shape[0] < Cn:
        bg = np.concatenate([bg, np.zeros(Cn - bg.shape[0], _f)])
    H, W = cam.H, cam.W
    color = np.zeros((Cn, H, W), _f); od = np.zeros((H, W), _f); oa = np.zeros((H, W), _f)
    fT = np.zeros((H, W), _f); nc = np.zeros((H, W), np.uint32); flags = np.zeros((H, W), np.uint8)
    pl = plist if N > 0 else np.zeros(1, np.uint32)
    L.ogs_oracle_blend_forward(C.c_int(W), C.c_int(H), C.c_int(Cn), _p(ranges), _p(pl), _p(xy),
                               _p(co), _p(colors), _p(depth), _p(bg), C.c_float(margin),
                               _p(color), _p(od), _p(oa), _p(fT), _p(nc), _p(flags))
    inputs = dict(means3D=_c(means3D), opacities=_c(opacities).reshape(-1), scales=_c(scales),
                  rotations=_c(rotations), cov3D_precomp=_c(cov3D_precomp), shs=_c(shs))
    return ForwardState(cam, P, Cn, radii, xy, depth, cov3D, co, rgb, clamped, tiles, offsets, N,
                        keys, plist, ranges, colors, bg, color, od, oa, fT, nc, flags, inputs)


def backward(st: ForwardState, dL_dcolor, dL_ddepth=None, dL_dalpha=None):
    """Returns dict of float64 gradients for one forward state."""
    L = lib()
    cam, P, Cn = st.cam, st.P, st.C
    H, W = cam.H, cam.W
    dL_dcolor = _c(dL_dcolor).reshape(Cn, H, W)
    dL_ddepth = _c(dL_ddepth); dL_dalpha = _c(dL_dalpha)
    d = np.float64
    g_m2 = np.zeros((P, 2), d); g_con = np.zeros((P, 3), d); g_op = np.zeros(P, d)
    g_col = np.zeros((P, Cn), d); g_dep = np.zeros(P, d)
    pl = st.point_list if st.N > 0 else np.zeros(1, np.uint32)
    L.ogs_oracle_blend_backward(
        C.c_int(P), C.c_int(W), C.c_int(H), C.c_int(Cn), _p(st.ranges), _p(pl), _p(st.xy),
        _p(st.conic_opacity), _p(st.colors), _p(st.depth), _p(st.bg), _p(st.final_T),
        _p(st.n_contrib), _p(dL_dcolor), _p(dL_ddepth), _p(dL_dalpha),
        _p(g_m2), _p(g_con), _p(g_op), _p(g_col), _p(g_dep))
    inp = st.inputs
    shs = inp["shs"]
    M = 0 if shs is None else shs.shape[1]
    g_mean = np.zeros((P, 3), d); g_sc = np.zeros((P, 3), d); g_rot = np.zeros((P, 4), d)
    g_cov = np.zeros((P, 6), d); g_sh = np.zeros((P, max(M, 1), 3), d)
    L.ogs_oracle_preprocess_backward(
        C.c_int(P), C.c_int(cam.sh_degree), C.c_int(M), C.c_int(Cn), _p(inp["means3D"]),
        _p(inp["scales"]), _p(inp["rotations"]), _p(inp["cov3D_precomp"]), _p(shs),
        C.c_float(cam.scale_modifier), _p(_c(cam.view).reshape(-1)), _p(_c(cam.proj).reshape(-1)),
        _p(_c(cam.campos)), C.c_int(W), C.c_int(H), C.c_float(cam.tanfovx), C.c_float(cam.tanfovy),
        _p(st.radii), _p(st.cov3D), _p(st.clamped), _p(g_m2), _p(g_con), _p(g_col), _p(g_dep),
        _p(g_mean), _p(g_sc), _p(g_rot), _p(g_cov), _p(g_sh) if shs is not None else None)
    out = dict(means3D=g_mean, means2D=np.concatenate([g_m2, np.zeros((P, 1), d)], 1),
               opacities=g_op.reshape(P, 1), scales=g_sc, rotations=g_rot, cov3D_precomp=g_cov,
               shs=g_sh if shs is not None else None, conic=g_con, depth=g_dep)
    if shs is None:
        out["colors_precomp"] = g_col[:, :3]
    out["extra"] = g_col[:, 3:] if Cn > 3 else None
    out["colors_all"] = g_col
    return out


def mark_visible(means3D, view):
    L = lib()
    P = means3D.shape[0]
    out = np.zeros(P, np.uint8)
    L.ogs_oracle_mark_visible(C.c_int(P), _p(_c(means3D)), _p(_c(view).reshape(-1)), _p(out))
    return out.astype(bool)
