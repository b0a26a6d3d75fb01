"""CPU restatement of the single-splat footprint vote of the SAM-mask refiner -- TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py).

Follows utils/sam_refinement_utils.py::MultiViewSAMMaskRefiner.get_splat_id_and_weights (:902-913):
render_single_gaussian (:330-403; one Gaussian, white view-independent SH, black background) -> fix_image
(:143-176: clamp(x * 255, 0, 255) -> uint8) -> rgb_to_weight_map (:103-141: channel mean / max) ->
get_most_common_id_in_mask_weighted (:645-702: ids shifted by -min, torch.bincount(weights), argmax).
The splat's geometry comes from the C oracle's preprocess (oracle/raster_oracle.c); a single Gaussian over a
black background leaves pixel = colour * alpha in the tiles its rectangle touches.

Pinned: tests/golden/make_footprint_golden.py renders the same splats one at a time with the C oracle's full
forward pass and pushes the images through the REFERENCE's own fix_image / rgb_to_weight_map /
get_most_common_id_in_mask_weighted (taken from the source with `ast`); tests/test_oracle_cpu.py requires this
restatement to return the same ids, visibility flags and footprint sizes.
"""
import numpy as np

from . import raster as orc

WHITE_SH_COLOR = np.float32(0.28209479177387814) + np.float32(0.5)


def splat_q_image(cam, radius, xy, conic_opacity, color=WHITE_SH_COLOR):
    """uint8 image [H,W] (all three channels are equal) of ONE splat; zeros where it does not reach."""
    f = np.float32
    H, W = cam.H, cam.W
    q = np.zeros((H, W), np.uint8)
    if radius <= 0:
        return q
    gx, gy = (W + 15) // 16, (H + 15) // 16
    r = f(radius)
    clampi = lambda v, hi: min(hi, max(0, int(v)))  # noqa: E731  (C cast: truncation toward zero)
    x0, y0 = clampi((f(xy[0]) - r) / f(16), gx), clampi((f(xy[1]) - r) / f(16), gy)
    x1, y1 = clampi((f(xy[0]) + r + f(15)) / f(16), gx), clampi((f(xy[1]) + r + f(15)) / f(16), gy)
    px0, px1, py0, py1 = x0 * 16, min(W, x1 * 16), y0 * 16, min(H, y1 * 16)
    if px1 <= px0 or py1 <= py0:
        return q
    xs = np.arange(px0, px1, dtype=f)[None, :]
    ys = np.arange(py0, py1, dtype=f)[:, None]
    dx, dy = f(xy[0]) - xs, f(xy[1]) - ys
    A, B, Cc, op = (f(v) for v in conic_opacity)
    power = f(-0.5) * (A * dx * dx + Cc * dy * dy) - B * dx * dy
    alpha = np.minimum(f(0.99), op * np.exp(power, dtype=f))
    ok = (power <= 0) & (alpha >= f(1.0 / 255.0))
    pix = np.where(ok, f(color) * alpha, f(0))
    q[py0:py1, px0:px1] = np.clip(pix * f(255), 0, 255).astype(np.uint8)      # fix_image :161-163
    return q


def vote(q, sam_ids):
    """(dominant id, its integer weight) -- rgb_to_weight_map + get_most_common_id_in_mask_weighted on integer
    weights (the normalisation by 255 * max is a positive factor); all-zero weights -> the smallest id."""
    ids = sam_ids.reshape(-1).astype(np.int64)
    lo, hi = int(ids.min()), int(ids.max())
    if lo == hi:
        return lo, int(q.sum())
    off = -lo if lo < 0 else 0
    counts = np.bincount(ids + off, weights=q.reshape(-1).astype(np.float64), minlength=hi + off + 1)
    k = int(np.argmax(counts))
    return k - off, int(counts[k])


def splat_votes(cam, means3D, opacities, scales, rotations, sam_ids):
    """Batched get_splat_id_and_weights for one camera: dict of [P] arrays (dominant_id, dominant_weight,
    footprint_pixels, q_max, radii, visible)."""
    radii, xy, _depth, _cov, co, _rgb, _cl, _tiles = orc.preprocess(cam, means3D, opacities, scales, rotations)
    P = means3D.shape[0]
    out = {k: np.zeros(P, np.int64) for k in ("dominant_id", "dominant_weight", "footprint_pixels", "q_max")}
    for i in range(P):
        q = splat_q_image(cam, int(radii[i]), xy[i], co[i])
        out["dominant_id"][i], out["dominant_weight"][i] = vote(q, sam_ids)
        out["footprint_pixels"][i] = int(np.count_nonzero(q))
        out["q_max"][i] = int(q.max())
    out["radii"] = radii.astype(np.int64)
    out["visible"] = out["footprint_pixels"] > 0
    return out
