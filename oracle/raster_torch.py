"""Independent fp64 autograd restatement of the rasterizer forward (TEST INFRASTRUCTURE ONLY).

Purpose: validate the hand-derived backward of oracle/raster_oracle.c.  The forward formulas
are written with differentiable torch ops; torch.autograd supplies the gradients.  Tile lists
(`point_list`, `ranges`) are taken from the C oracle's integer binning.  Two upstream
conventions are reproduced explicitly because plain autograd would differ:
  * the 0.99 alpha clamp is straight-through (the CUDA backward ignores it);
  * when t.x/t.z is clamped to +-1.3 tan(fov) the clamped value is treated as a constant.
Small scenes only (python loop over tiles).
"""
import numpy as np
import torch

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792,
      0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
      -0.4570457994644658, 1.445305721320277, -0.5900435899266435]


def _sh_rgb(deg, sh, dirs):
    """utils/sh_utils.py:57-112 basis; sh [P, M, 3], dirs [P, 3] -> [P, 3]."""
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    res = C0 * sh[:, 0]
    if deg > 0:
        res = res - C1 * y * sh[:, 1] + C1 * z * sh[:, 2] - C1 * x * sh[:, 3]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        res = (res + C2[0] * xy * sh[:, 4] + C2[1] * yz * sh[:, 5] + C2[2] * (2 * zz - xx - yy) * sh[:, 6]
               + C2[3] * xz * sh[:, 7] + C2[4] * (xx - yy) * sh[:, 8])
    if deg > 2:
        res = (res + C3[0] * y * (3 * xx - yy) * sh[:, 9] + C3[1] * xy * z * sh[:, 10]
               + C3[2] * y * (4 * zz - xx - yy) * sh[:, 11] + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12]
               + C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] + C3[5] * z * (xx - yy) * sh[:, 14]
               + C3[6] * x * (xx - 3 * yy) * sh[:, 15])
    return res


def render(cam, radii, point_list, ranges, means3D, means2D, opacities, scales=None, rotations=None,
           cov3D_precomp=None, shs=None, colors_precomp=None, extra=None, bg=None):
    """All tensor args are float64 leaves (requires_grad as the caller wishes).
    Returns color [C,H,W], depth [H,W], alpha [H,W]."""
    dt = torch.float64
    W, H = cam.W, cam.H
    view = torch.as_tensor(np.asarray(cam.view, dtype=np.float64).reshape(4, 4), dtype=dt)
    proj = torch.as_tensor(np.asarray(cam.proj, dtype=np.float64).reshape(4, 4), dtype=dt)
    campos = torch.as_tensor(np.asarray(cam.campos, dtype=np.float64), dtype=dt)
    P = means3D.shape[0]
    ones = torch.ones(P, 1, dtype=dt)
    hom = torch.cat([means3D, ones], 1)
    p_view = hom @ view            # row-vector convention
    p_hom = hom @ proj
    p_w = 1.0 / (p_hom[:, 3] + 1e-7)
    ndc = p_hom[:, :2] * p_w[:, None]
    xy = torch.stack([((ndc[:, 0] + 1) * W - 1) * 0.5, ((ndc[:, 1] + 1) * H - 1) * 0.5], 1)
    xy = xy + means2D[:, :2] * torch.tensor([0.5 * W, 0.5 * H], dtype=dt)
    if cov3D_precomp is None:
        q = rotations
        r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        R = torch.stack([
            1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
            2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
            2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).reshape(P, 3, 3)
        L = R * (cam.scale_modifier * scales)[:, None, :]
        Sigma = L @ L.transpose(1, 2)
    else:
        c = cov3D_precomp
        Sigma = torch.stack([c[:, 0], c[:, 1], c[:, 2], c[:, 1], c[:, 3], c[:, 4], c[:, 2], c[:, 4], c[:, 5]], 1).reshape(P, 3, 3)
    fx = float(np.float32(W) / (np.float32(2.0) * np.float32(cam.tanfovx)))
    fy = float(np.float32(H) / (np.float32(2.0) * np.float32(cam.tanfovy)))
    limx, limy = float(np.float32(1.3) * np.float32(cam.tanfovx)), float(np.float32(1.3) * np.float32(cam.tanfovy))
    tz = p_view[:, 2]
    txtz, tytz = p_view[:, 0] / tz, p_view[:, 1] / tz
    cx = (txtz < -limx) | (txtz > limx)
    cy = (tytz < -limy) | (tytz > limy)
    tx = torch.where(cx, (txtz.clamp(-limx, limx) * tz).detach(), p_view[:, 0])
    ty = torch.where(cy, (tytz.clamp(-limy, limy) * tz).detach(), p_view[:, 1])
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -(fx * tx) / (tz * tz), zero, fy / tz, -(fy * ty) / (tz * tz)], 1).reshape(P, 2, 3)
    Rw = view[:3, :3].t()          # standard world->view rotation
    Tm = J @ Rw
    cov2 = Tm @ Sigma @ Tm.transpose(1, 2)
    a = cov2[:, 0, 0] + 0.3
    b = cov2[:, 0, 1]
    c_ = cov2[:, 1, 1] + 0.3
    det = a * c_ - b * b
    det = torch.where(det == 0, torch.ones_like(det), det)
    conic = torch.stack([c_ / det, -b / det, a / det], 1)
    depth = p_view[:, 2]
    if shs is not None:
        d = means3D - campos
        d = d / d.norm(dim=1, keepdim=True)
        base = torch.clamp_min(_sh_rgb(cam.sh_degree, shs, d) + 0.5, 0.0)
    else:
        base = colors_precomp
    colors = base if extra is None else torch.cat([base, extra], 1)
    Cn = colors.shape[1]
    bgv = torch.zeros(Cn, dtype=dt) if bg is None else torch.as_tensor(np.asarray(bg, dtype=np.float64), dtype=dt)
    if bgv.shape[0] < Cn:
        bgv = torch.cat([bgv, torch.zeros(Cn - bgv.shape[0], dtype=dt)])
    op = opacities.reshape(-1)

    out_c = torch.zeros(Cn, H, W, dtype=dt)
    out_d = torch.zeros(H, W, dtype=dt)
    out_a = torch.zeros(H, W, dtype=dt)
    gx = (W + 15) // 16
    gy = (H + 15) // 16
    pl = torch.as_tensor(point_list.astype(np.int64))
    col_tiles, d_tiles, a_tiles = {}, {}, {}
    for tyi in range(gy):
        for txi in range(gx):
            t = tyi * gx + txi
            lo, hi = int(ranges[t, 0]), int(ranges[t, 1])
            y0, y1 = tyi * 16, min(tyi * 16 + 16, H)
            x0, x1 = txi * 16, min(txi * 16 + 16, W)
            ys, xs = torch.meshgrid(torch.arange(y0, y1, dtype=dt), torch.arange(x0, x1, dtype=dt), indexing="ij")
            npx = ys.numel()
            if hi <= lo:
                col_tiles[t] = bgv[:, None].expand(Cn, npx).reshape(Cn, y1 - y0, x1 - x0)
                continue
            ids = pl[lo:hi]
            dx = xy[ids, 0][None, :] - xs.reshape(-1, 1)
            dy = xy[ids, 1][None, :] - ys.reshape(-1, 1)
            cn = conic[ids]
            power = -0.5 * (cn[:, 0][None] * dx * dx + cn[:, 2][None] * dy * dy) - cn[:, 1][None] * dx * dy
            G = torch.exp(power)
            alpha_raw = op[ids][None] * G
            alpha = alpha_raw + (alpha_raw.clamp(max=0.99) - alpha_raw).detach()
            cand = (power <= 0) & (alpha >= 1.0 / 255.0)
            a_c = torch.where(cand, alpha, torch.zeros_like(alpha))
            one_m = 1 - a_c
            T_incl = torch.cumprod(one_m, dim=1)
            T_before = torch.cat([torch.ones(npx, 1, dtype=dt), T_incl[:, :-1]], 1)
            trig = cand & (T_incl < 1e-4)
            done = torch.cummax(trig.to(torch.int8), dim=1)[0].bool()
            keep = cand & ~done
            a_k = torch.where(keep, alpha, torch.zeros_like(alpha))
            w = a_k * T_before
            T_final = torch.prod(1 - a_k, dim=1)
            cpix = w @ colors[ids] + T_final[:, None] * bgv[None]
            out_c[:, y0:y1, x0:x1] = cpix.t().reshape(Cn, y1 - y0, x1 - x0)
            out_d[y0:y1, x0:x1] = (w @ depth[ids]).reshape(y1 - y0, x1 - x0)
            out_a[y0:y1, x0:x1] = w.sum(1).reshape(y1 - y0, x1 - x0)
    for t, v in col_tiles.items():
        tyi, txi = divmod(t, gx)
        y0, y1 = tyi * 16, min(tyi * 16 + 16, H)
        x0, x1 = txi * 16, min(txi * 16 + 16, W)
        out_c[:, y0:y1, x0:x1] = v
    return out_c, out_d, out_a
