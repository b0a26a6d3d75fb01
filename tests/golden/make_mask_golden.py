"""Generates tests/golden/mask_stats_golden.npz by running the REFERENCE's own code (torch CPU):

* utils/opengs_utlis.py::mask_feature_mean (:240-283), pair_mask_feature_mean (:184-201) -- the
  module is loaded by path with a stub for the missing `bitarray` package;
* utils/opengs_utlis.py::calculate_iou (:90-123) on the masks of one case against a second, shifted set, for
  base = None / "former" / "later" (`iou_inputs`);
* train.py::cohesion_loss (:102-121) and separation_loss (:123-155) -- train.py itself cannot be
  imported here (pytorch3d, plyfile, a CUDA device ...), so the two function definitions are taken
  from its source with `ast` and executed unchanged.

* utils/opengs_utlis.py::get_SAM_mask_and_feat (:125-182) on a synthetic 4-level SAM id map (`sam_inputs`), at levels 0
  and 3; the masks it returns (a PARTITION of the image, 20+ masks: the id-map path of csrc/mask_stats.cu) then go
  through the same Stage-1 loss.

Stored: the reference outputs and the autograd gradients of the Stage-1 loss
(train.py:450-456: loss = separation + 0.1 * cohesion) w.r.t. feat_map and image_mask.
Run in the build container only:  python tests/golden/make_mask_golden.py
"""
import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

CASES = {
    # name: (C, H, W, num_mask, with_image_mask)
    "stage1_6ch": (6, 47, 66, 9, True),         # H*W not a multiple of 4: unaligned mask rows
    "no_image_mask": (6, 32, 40, 5, False),
    "rgb_3ch": (3, 24, 28, 4, True),
}


def inputs(name):
    C, H, W, M, with_img = CASES[name]
    rs = np.random.RandomState({"stage1_6ch": 31, "no_image_mask": 32, "rgb_3ch": 33}[name])
    feat = rs.rand(C, H, W).astype(np.float32)
    # blocky, partly overlapping masks + one empty mask
    masks = np.zeros((M, H, W), bool)
    for m in range(M - 1):
        y0, x0 = rs.randint(0, H - 6), rs.randint(0, W - 6)
        h, w = rs.randint(4, H // 2), rs.randint(4, W // 2)
        masks[m, y0:y0 + h, x0:x0 + w] = rs.rand(min(h, H - y0), min(w, W - x0)) > 0.2
    img = (rs.rand(1, H, W) > 0.15).astype(np.float32) * rs.rand(1, H, W).astype(np.float32) if with_img else None
    return feat, masks, img


def iou_inputs(name):
    """Second mask set for calculate_iou: the case's masks rolled by a few pixels, two of them merged, plus an
    empty one and a full one; returned as (masks1 [n,H,W], masks2 [m,H,W])."""
    _, masks, _ = inputs(name)
    other = np.roll(masks[: max(2, masks.shape[0] // 2)], (2, -3), axis=(1, 2)).copy()
    other[0] |= other[1]
    other = np.concatenate([other, np.zeros_like(masks[:1]), np.ones_like(masks[:1])])
    return masks, other


def sam_inputs():
    """A 4-level SAM id map [4,H,W] (int32) as the dataset stores it: level l's ids continue after level l-1's maximum,
    -1 marks pixels without a mask; plus a feature map and a silhouette for the Stage-1 loss."""
    H, W = 40, 52
    rs = np.random.RandomState(41)
    levels, base = [], 0
    for l, (gy, gx) in enumerate(((3, 4), (4, 5), (2, 3), (5, 5))):
        yy, xx = np.mgrid[0:H, 0:W]
        ids = (yy * gy // H) * gx + (xx * gx // W) + base
        ids = np.where(rs.rand(H, W) < 0.07, -1, ids)                 # invalid pixels
        ids[yy + xx * (l + 1) % 7 == 3] = -1
        levels.append(ids.astype(np.int32))
        base = int(ids.max()) + 1
    feat = rs.rand(6, H, W).astype(np.float32)
    img = ((rs.rand(1, H, W) > 0.1) * rs.rand(1, H, W)).astype(np.float32)
    return np.stack(levels), feat, img


def load_reference():
    sys.modules.setdefault("bitarray", types.SimpleNamespace(bitarray=object))
    spec = importlib.util.spec_from_file_location("ref_opengs_utlis", os.path.join(REF, "utils/opengs_utlis.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    src = open(os.path.join(REF, "train.py")).read()
    tree = ast.parse(src)
    ns = {"torch": torch}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("cohesion_loss", "separation_loss"):
            exec(compile(ast.Module(body=[node], type_ignores=[]), "train.py", "exec"), ns)
    return mod, ns["cohesion_loss"], ns["separation_loss"]


def main():
    ref, cohesion_loss, separation_loss = load_reference()
    out = {}
    for name in CASES:
        feat_np, masks_np, img_np = inputs(name)
        feat = torch.from_numpy(feat_np).requires_grad_(True)
        masks = torch.from_numpy(masks_np)
        img = None if img_np is None else torch.from_numpy(img_np).requires_grad_(True)
        mean = ref.mask_feature_mean(feat, masks, image_mask=img)
        loss_c = cohesion_loss(feat, masks, mean)
        loss_s = separation_loss(mean, 1000)
        loss = loss_s + 0.1 * loss_c
        loss.backward()
        out[f"{name}/mean"] = mean.detach().numpy()
        out[f"{name}/cohesion"] = loss_c.detach().numpy()
        out[f"{name}/separation"] = loss_s.detach().numpy()
        out[f"{name}/dfeat"] = feat.grad.numpy()
        if img is not None:
            out[f"{name}/dimg"] = img.grad.numpy()
        with torch.no_grad():
            m2, var, cnt = ref.mask_feature_mean(feat.detach(), masks, return_var=True)
            out[f"{name}/mean_noimg"] = m2.numpy()
            out[f"{name}/var"] = var.numpy()
            out[f"{name}/cnt"] = cnt.numpy()
            pm = ref.pair_mask_feature_mean(feat.detach().unsqueeze(0).repeat(3, 1, 1, 1), masks[:3])
            out[f"{name}/pair_mean"] = pm.numpy()
            m1, m2 = iou_inputs(name)
            for base in (None, "former", "later"):
                # int32 inputs exercise the reference's .to(torch.bool) branch (:100-103)
                iou = ref.calculate_iou(torch.from_numpy(m1), torch.from_numpy(m2.astype(np.int32)), base=base)
                out[f"{name}/iou_{base}"] = iou.numpy()
    sam, feat_np, img_np = sam_inputs()
    for level in (0, 3):
        mask_id, mask_bool, invalid = ref.get_SAM_mask_and_feat(torch.from_numpy(sam), level=level)
        feat = torch.from_numpy(feat_np).requires_grad_(True)
        img = torch.from_numpy(img_np).requires_grad_(True)
        mean = ref.mask_feature_mean(feat, mask_bool, image_mask=img)
        loss_c = cohesion_loss(feat, mask_bool, mean)
        loss_s = separation_loss(mean, 1000)
        (loss_s + 0.1 * loss_c).backward()
        k = f"sam_l{level}"
        out[f"{k}/mask_id"] = mask_id.numpy()
        out[f"{k}/mask_bool"] = np.packbits(mask_bool.numpy().astype(bool), axis=None)
        out[f"{k}/mask_bool_shape"] = np.array(mask_bool.shape)
        out[f"{k}/invalid_pix"] = invalid.numpy()
        out[f"{k}/mean"] = mean.detach().numpy()
        out[f"{k}/cohesion"] = loss_c.detach().numpy()
        out[f"{k}/separation"] = loss_s.detach().numpy()
        out[f"{k}/dfeat"] = feat.grad.numpy()
        out[f"{k}/dimg"] = img.grad.numpy()
        with torch.no_grad():
            _, var, cnt = ref.mask_feature_mean(feat.detach(), mask_bool, return_var=True)
            out[f"{k}/var"] = var.numpy()
            out[f"{k}/cnt"] = cnt.numpy()
    path = os.path.join(HERE, "mask_stats_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, sorted(out)[:8])


if __name__ == "__main__":
    main()
