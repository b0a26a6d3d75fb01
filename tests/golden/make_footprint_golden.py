"""Generates tests/golden/footprint_golden.npz: the REFERENCE's own post-processing of a single-splat render --
fix_image (:143-176), rgb_to_weight_map (:103-141) and MultiViewSAMMaskRefiner.get_most_common_id_in_mask_weighted
(:645-702) of utils/sam_refinement_utils.py, taken from the source with `ast` (the module imports rerun, cv2,
the CUDA rasterizer ... and cannot be imported) -- applied to one-Gaussian images rendered with the C oracle's
FULL forward pass (oracle/raster.py::forward, white view-independent SH, black background), which stands in for
render_single_gaussian (:330-403; the CUDA rasterizer is an un-vendored dependency).

Stored per splat: dominant id, render-visible flag (`non_black_mask.any()`), non-black pixel count, uint8 maximum.
Run in the build container only:  python tests/golden/make_footprint_golden.py
"""
import ast
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference/utils/sam_refinement_utils.py"

CASES = {"ids_from_minus1": (70, 96, 72, 0, -1), "ids_from_3": (50, 80, 80, 1, 3)}   # P, W, H, seed, lowest id


def inputs(name):
    import helpers
    P, W, H, seed, lo = CASES[name]
    gs, cam = helpers.small_scene(P=P, W=W, H=H, seed=seed, scale_mult=0.5)
    # SAM-like id map: 12 x 10 pixel blocks with shuffled ids, so that a footprint straddles several segments
    rs = np.random.RandomState(seed + 10)
    nby, nbx = (H + 9) // 10, (W + 11) // 12
    block_ids = rs.permutation(nby * nbx).reshape(nby, nbx) % 23
    sam = np.repeat(np.repeat(block_ids, 10, axis=0), 12, axis=1)[:H, :W].astype(np.int32) + lo
    return gs, cam, sam


def load_reference():
    tree = ast.parse(open(REF).read())
    ns = {"torch": torch, "np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("fix_image", "rgb_to_weight_map"):
            exec(compile(ast.Module(body=[node], type_ignores=[]), REF, "exec"), ns)
        if isinstance(node, ast.ClassDef) and node.name == "MultiViewSAMMaskRefiner":
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name == "get_most_common_id_in_mask_weighted":
                    exec(compile(ast.Module(body=[item], type_ignores=[]), REF, "exec"), ns)
    return ns["fix_image"], ns["rgb_to_weight_map"], ns["get_most_common_id_in_mask_weighted"]


def main():
    import helpers
    from oracle import raster as orc
    fix_image, rgb_to_weight_map, most_common = load_reference()
    me = types.SimpleNamespace(verbose_logging=False)
    out = {}
    for name in CASES:
        gs, cam, sam = inputs(name)
        ocam = helpers.to_oracle_cam(cam)
        g = helpers.np_inputs(gs)
        P = g["means3D"].shape[0]
        white = np.zeros((1, 16, 3), np.float32)
        white[:, 0, :] = 1.0
        dom, vis, npx, qmx = [], [], [], []
        for i in range(P):
            sl = slice(i, i + 1)
            st = orc.forward(ocam, g["means3D"][sl], g["opacities"][sl], g["scales"][sl], g["rotations"][sl], shs=white,
                             bg=np.zeros(3, np.float32))
            img = fix_image(torch.from_numpy(st.color[:3].copy()))          # [H,W,3] uint8, reference :908
            non_black = torch.any(img != 0, dim=2)                           # :909
            weights = rgb_to_weight_map(img)                                 # :910
            dom.append(most_common(me, sam_mask=torch.from_numpy(sam), weight_matrix=weights))   # :911
            vis.append(bool(non_black.any()))
            npx.append(int(non_black.sum()))
            qmx.append(int(img.max()))
        out[f"{name}/dominant_id"] = np.array(dom, np.int64)
        out[f"{name}/visible"] = np.array(vis)
        out[f"{name}/footprint_pixels"] = np.array(npx, np.int64)
        out[f"{name}/q_max"] = np.array(qmx, np.int64)
        print(name, "visible", int(np.sum(vis)), "of", P, "ids", sorted(set(dom)))
    path = os.path.join(HERE, "footprint_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
