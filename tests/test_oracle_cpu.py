"""CPU tests (no GPU): the oracles against (i) the golden vectors produced by the reference
k-means file itself, (ii) an independent fp64 autograd restatement, (iii) structural properties."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from helpers import np_inputs, small_scene, to_oracle_cam
from oracle import kmeans as okm
from oracle import raster as orc
from oracle import raster_torch as rt

GOLD = os.path.join(os.path.dirname(__file__), "golden", "kmeans_golden.npz")


def _golden_inputs(name):
    spec = importlib.util.spec_from_file_location("mkg", os.path.join(os.path.dirname(GOLD), "make_kmeans_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.inputs(name), m.CASES[name]


@pytest.mark.parametrize("name", ["root_25k", "root_20k_exact_chunks"])
def test_kmeans_oracle_vs_reference_golden_root(name):
    gold = np.load(GOLD)
    (ins_feat, xyz), (N, k1, k2, iters, pw) = _golden_inputs(name)
    c0 = np.concatenate([ins_feat, xyz * np.float32(pw)], 1)[:k1]
    centers, ids = okm.cluster_assign_root(ins_feat, xyz, pw, c0, iters)
    ref_ids = gold[f"{name}/cls_ids"].astype(np.int64)
    mism = (ids != ref_ids).mean()
    # the reference's cdist uses the cancellation-prone matmul form with an unspecified BLAS order:
    # agreement is required away from near-ties only ("k-means bit-exact ids")
    assert mism <= 2e-3, mism
    assert np.allclose(centers, gold[f"{name}/centers"], rtol=2e-3, atol=2e-3)
    q = centers[ids[:512]][:, :6]
    assert np.abs(q - gold[f"{name}/ins_feat_q"]).max() < 0.25   # a few near-tie points may switch centre
    assert (np.abs(q - gold[f"{name}/ins_feat_q"]).max(1) > 1e-3).mean() <= 0.01


def test_kmeans_oracle_vs_reference_golden_leaf():
    gold = np.load(GOLD)
    name = "root_25k"
    (ins_feat, xyz), (N, k1, k2, iters, pw) = _golden_inputs(name)
    cls_ids = gold[f"{name}/cls_ids"].astype(np.int64)          # coarse ids as the reference produced them
    leaf_centers = ins_feat[:k1 * k2 + 1].copy()
    leaf_ids = np.full(N, k1 * k2, np.int64)
    for sel, n_sub in ((3, k2), (7, 4)):
        leaf_centers, leaf_ids = okm.cluster_assign_leaf(ins_feat, cls_ids, leaf_centers, leaf_ids, sel, n_sub, k2, iters)
    ref_ids = gold[f"{name}/leaf_cls_ids"].astype(np.int64)
    assert (leaf_ids != ref_ids).mean() <= 2e-3
    assert np.array_equal(leaf_ids == k1 * k2, ref_ids == k1 * k2)      # sentinel for untouched points
    assert np.allclose(leaf_centers, gold[f"{name}/leaf_centers"], rtol=2e-3, atol=2e-3)
    # rows >= iLeafSubNum of a rewritten block are zero (0 / eps), rows of untouched blocks keep their init
    assert np.all(leaf_centers[7 * k2 + 4:8 * k2] == 0.0)
    assert np.array_equal(leaf_centers[0:k2], ins_feat[0:k2])


def test_kmeans_assign_tie_and_select():
    a = np.array([[0.0, 0.0], [1.0, 1.0], [0.5, 0.5]], np.float32)
    c = np.array([[1.0, 1.0], [0.0, 0.0], [1.0, 1.0]], np.float32)
    ids = okm.assign(a, None, 1.0, c)
    assert ids.tolist() == [1, 0, 0]          # exact tie at [0.5,0.5] -> lowest index; duplicate centre -> first
    sel = np.array([5, 6, 5], np.int64)
    out = np.full(3, -7, np.int64)
    okm.assign(a, None, 1.0, c, sel, 5, 100, out)
    assert out.tolist() == [101, -7, 100]


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-20)


@pytest.mark.parametrize("mode", ["sh", "cov_precomp_colors"])
def test_raster_oracle_backward_vs_fp64_autograd(mode):
    P, W, H = 150, 48, 32
    gs, cam = small_scene(P=P, W=W, H=H, seed=1)
    ocam = to_oracle_cam(cam)
    g = np_inputs(gs)
    bg = np.array([0.1, 0.2, 0.3], np.float32)
    rng = np.random.default_rng(0)
    colors = rng.random((P, 3)).astype(np.float32)
    if mode == "sh":
        st = orc.forward(ocam, g["means3D"], g["opacities"], g["scales"], g["rotations"], shs=g["shs"],
                         extra=g["ins_feat"], bg=bg)
    else:
        st0 = orc.forward(ocam, g["means3D"], g["opacities"], g["scales"], g["rotations"], shs=g["shs"], bg=bg)
        st = orc.forward(ocam, g["means3D"], g["opacities"], cov3D_precomp=st0.cov3D * 0 + _cov(g), colors_precomp=colors,
                         extra=g["ins_feat"], bg=bg)
    gc = rng.standard_normal((9, H, W)).astype(np.float32)
    gd = rng.standard_normal((H, W)).astype(np.float32)
    ga = rng.standard_normal((H, W)).astype(np.float32)
    gr = orc.backward(st, gc, gd, ga)

    T = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), requires_grad=True)  # noqa: E731
    m3, op, ex = T(g["means3D"]), T(g["opacities"]), T(g["ins_feat"])
    m2 = torch.zeros(P, 3, dtype=torch.float64, requires_grad=True)
    kw, leaves = {}, dict(means3D=m3, means2D=m2, opacities=op, extra=ex)
    if mode == "sh":
        leaves.update(scales=T(g["scales"]), rotations=T(g["rotations"]), shs=T(g["shs"]))
        kw = dict(scales=leaves["scales"], rotations=leaves["rotations"], shs=leaves["shs"])
    else:
        leaves.update(cov3D_precomp=T(_cov(g)), colors_precomp=T(colors))
        kw = dict(cov3D_precomp=leaves["cov3D_precomp"], colors_precomp=leaves["colors_precomp"])
    c, d, a = rt.render(ocam, st.radii, st.point_list, st.ranges, m3, m2, op, extra=ex, bg=bg.astype(np.float64), **kw)
    assert np.abs(c.detach().numpy() - st.color).max() < 5e-6
    assert np.abs(a.detach().numpy() - st.out_alpha).max() < 5e-6
    L = (c * torch.tensor(gc, dtype=torch.float64)).sum() + (d * torch.tensor(gd, dtype=torch.float64)).sum() + \
        (a * torch.tensor(ga, dtype=torch.float64)).sum()
    L.backward()
    for k, t in leaves.items():
        assert _rel(gr[k], t.grad.numpy().reshape(np.asarray(gr[k]).shape)) < 2e-5, k


def _cov(g):
    q = g["rotations"].astype(np.float64)
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                  2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                  2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).reshape(-1, 3, 3)
    L = R * g["scales"].astype(np.float64)[:, None, :]
    S = L @ L.transpose(0, 2, 1)
    return np.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], 1).astype(np.float32)


def test_raster_oracle_structure():
    gs, cam = small_scene(P=2000, W=100, H=70, seed=2, scale_mult=3.0)
    g = np_inputs(gs)
    st = orc.forward(to_oracle_cam(cam), g["means3D"], g["opacities"], g["scales"], g["rotations"], shs=g["shs"])
    assert st.N == int(st.tiles_touched.sum()) == len(st.keys)
    assert np.all(st.keys[1:] >= st.keys[:-1])
    tiles = (st.keys >> np.uint64(32)).astype(np.int64)
    for t in np.unique(tiles)[:50]:
        lo, hi = st.ranges[t]
        assert np.all(tiles[lo:hi] == t) and (lo == 0 or tiles[lo - 1] != t) and (hi == st.N or tiles[hi] != t)
    # depth order inside a tile, ties broken by Gaussian index (stable sort of emission order)
    lo, hi = st.ranges[np.bincount(tiles).argmax()]
    d = st.depth[st.point_list[lo:hi]]
    assert np.all(np.diff(d) >= 0)
    assert st.out_alpha.min() >= 0 and st.out_alpha.max() <= 1 + 1e-6
    assert np.allclose(st.out_alpha, 1 - st.final_T, atol=2e-6)
    # empty scene -> background
    st0 = orc.forward(to_oracle_cam(cam), g["means3D"][:0], g["opacities"][:0], g["scales"][:0], g["rotations"][:0],
                      shs=g["shs"][:0], bg=np.array([0.2, 0.4, 0.6], np.float32))
    assert st0.N == 0 and np.allclose(st0.color[1], 0.4)


# ---- the reference's own Python code for pieces of the rasterizer path (tests/golden/make_raster_golden.py) ----
RGOLD = os.path.join(os.path.dirname(__file__), "golden", "raster_pieces_golden.npz")


def _raster_golden_inputs():
    spec = importlib.util.spec_from_file_location("mrg", os.path.join(os.path.dirname(RGOLD), "make_raster_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.inputs()


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_raster_oracle_sh_colours_vs_reference_eval_sh(deg):
    """oracle SH -> RGB (+0.5, clamp at 0, clamp mask) == utils/sh_utils.py::eval_sh as used at
    gaussian_renderer/__init__.py:90-96, on the committed golden vectors."""
    from opengaussian_b200 import synth
    inp, gold = _raster_golden_inputs(), np.load(RGOLD)
    cam = synth.look_at(tuple(float(v) for v in inp["campos"]), (0.0, 0.0, 0.0), 640, 480, 1.3)
    oc = to_oracle_cam(cam, sh_degree=deg)
    P = inp["xyz"].shape[0]
    radii, xy, depth, cov3D, co, rgb, clamped, tiles = orc.preprocess(
        oc, inp["xyz"], np.full((P, 1), 0.5, np.float32), inp["scales"], inp["rot"], None, inp["shs"])
    vis = radii > 0
    assert vis.mean() > 0.8
    assert np.abs(rgb[vis] - gold[f"rgb_deg{deg}"][vis]).max() <= 2e-6
    unclamped = gold[f"rgb_unclamped_deg{deg}"][vis]
    sure = np.abs(unclamped) > 1e-5                       # away from the clamp threshold
    assert np.array_equal((clamped[vis] != 0)[sure], (unclamped < 0)[sure])


@pytest.mark.parametrize("mod", [1.0, 0.7])
def test_raster_oracle_cov3d_vs_reference_build_covariance(mod):
    """oracle cov3D (scale_modifier * scale, quaternion (r,x,y,z), packing xx,xy,xz,yy,yz,zz) ==
    utils/general_utils.py build_scaling_rotation / strip_symmetric (scene/gaussian_model.py:63-67)."""
    from opengaussian_b200 import synth
    inp, gold = _raster_golden_inputs(), np.load(RGOLD)
    cam = synth.look_at(tuple(float(v) for v in inp["campos"]), (0.0, 0.0, 0.0), 640, 480, 1.3)
    oc = to_oracle_cam(cam, scale_modifier=mod)
    P = inp["xyz"].shape[0]
    radii, xy, depth, cov3D, co, rgb, clamped, tiles = orc.preprocess(
        oc, inp["xyz"], np.full((P, 1), 0.5, np.float32), inp["scales"], inp["rot"], None, inp["shs"])
    vis = radii > 0
    want = gold[f"cov3D_mod{mod}"][vis]
    assert np.abs(cov3D[vis] - want).max() <= 1e-6 * np.abs(want).max() + 1e-9


def test_camera_matrices_vs_reference_graphics_utils():
    """synth.camera_from_RT == scene/cameras.py:71-78 built from utils/graphics_utils.py (golden)."""
    from opengaussian_b200 import synth
    inp, gold = _raster_golden_inputs(), np.load(RGOLD)
    for i, (R, T, fx, fy) in enumerate(inp["cams"]):
        c = synth.camera_from_RT(R, T, fx, fy, 640, 480)
        assert np.abs(c.world_view_transform.numpy() - gold[f"cam{i}_world_view"]).max() <= 1e-6
        assert np.abs(c.full_proj_transform.numpy() - gold[f"cam{i}_full_proj"]).max() <= 2e-5
        assert np.abs(c.camera_center.numpy() - gold[f"cam{i}_center"]).max() <= 1e-5
        assert abs(c.tanfovx - gold[f"cam{i}_tan"][0]) <= 1e-12 and abs(c.tanfovy - gold[f"cam{i}_tan"][1]) <= 1e-12
        assert np.abs(synth.projection_matrix(fx, fy).t().float().numpy() - gold[f"cam{i}_proj"]).max() <= 1e-6


# ---- per-mask statistics / Stage-1 losses: oracle vs the reference's own functions (golden) ----
MGOLD = os.path.join(os.path.dirname(__file__), "golden", "mask_stats_golden.npz")


def _mask_golden_module():
    spec = importlib.util.spec_from_file_location("mmg", os.path.join(os.path.dirname(MGOLD), "make_mask_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("name", ["stage1_6ch", "no_image_mask", "rgb_3ch"])
def test_mask_stats_oracle_vs_reference_golden(name):
    """oracle/mask_stats.py == utils/opengs_utlis.py::mask_feature_mean + train.py::cohesion_loss /
    separation_loss run on the CPU (values and autograd gradients of the Stage-1 loss)."""
    from oracle import mask_stats as oms
    m, gold = _mask_golden_module(), np.load(MGOLD)
    feat_np, masks_np, img_np = m.inputs(name)
    feat = torch.from_numpy(feat_np).requires_grad_(True)
    masks = torch.from_numpy(masks_np)
    img = None if img_np is None else torch.from_numpy(img_np).requires_grad_(True)
    mean = oms.mask_feature_mean(feat, masks, image_mask=img)
    lc = oms.cohesion_loss(feat, masks, mean)
    ls = oms.separation_loss(mean, 1000)
    (ls + 0.1 * lc).backward()
    assert np.abs(mean.detach().numpy() - gold[f"{name}/mean"]).max() <= 2e-6
    assert abs(float(lc) - float(gold[f"{name}/cohesion"])) <= 2e-6
    assert abs(float(ls) - float(gold[f"{name}/separation"])) <= 2e-6
    gd = gold[f"{name}/dfeat"]
    assert np.abs(feat.grad.numpy() - gd).max() <= 1e-5 * np.abs(gd).max() + 1e-9
    if img is not None:
        gi = gold[f"{name}/dimg"]
        assert np.abs(img.grad.numpy() - gi).max() <= 1e-5 * np.abs(gi).max() + 1e-9
    m2, var, cnt = oms.mask_feature_mean(feat.detach(), masks, return_var=True)
    assert np.abs(m2.numpy() - gold[f"{name}/mean_noimg"]).max() <= 2e-6
    assert np.abs(var.numpy() - gold[f"{name}/var"]).max() <= 2e-6
    assert np.array_equal(cnt.numpy(), gold[f"{name}/cnt"])


@pytest.mark.parametrize("name", ["stage1_6ch", "no_image_mask", "rgb_3ch"])
def test_iou_oracle_vs_reference_golden(name):
    """oracle calculate_iou == utils/opengs_utlis.py::calculate_iou (:90-123), bit for bit."""
    from oracle import mask_stats as oms
    m, gold = _mask_golden_module(), np.load(MGOLD)
    m1, m2 = m.iou_inputs(name)
    for base in (None, "former", "later"):
        iou = oms.calculate_iou(torch.from_numpy(m1), torch.from_numpy(m2.astype(np.int32)), base=base)
        ref = gold[f"{name}/iou_{base}"]
        assert iou.shape == ref.shape == (m2.shape[0], m1.shape[0])
        assert np.array_equal(iou.numpy(), ref), base


# ---- single-splat footprint votes: oracle vs the reference's own post-processing (golden) ----
FGOLD = os.path.join(os.path.dirname(__file__), "golden", "footprint_golden.npz")


def _footprint_golden_module():
    spec = importlib.util.spec_from_file_location("mfg", os.path.join(os.path.dirname(FGOLD), "make_footprint_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("name", ["ids_from_minus1", "ids_from_3"])
def test_footprint_oracle_vs_reference_golden(name):
    """oracle/footprint.py (rect-restricted single-splat image + integer vote) == the C oracle's full P = 1 render
    pushed through the reference's fix_image / rgb_to_weight_map / get_most_common_id_in_mask_weighted."""
    import helpers
    from oracle import footprint as ofp
    m, gold = _footprint_golden_module(), np.load(FGOLD)
    gs, cam, sam = m.inputs(name)
    g = helpers.np_inputs(gs)
    out = ofp.splat_votes(helpers.to_oracle_cam(cam), g["means3D"], g["opacities"], g["scales"], g["rotations"], sam)
    assert np.array_equal(out["visible"], gold[f"{name}/visible"])
    assert np.array_equal(out["footprint_pixels"], gold[f"{name}/footprint_pixels"])
    assert np.array_equal(out["q_max"], gold[f"{name}/q_max"])
    assert np.array_equal(out["dominant_id"], gold[f"{name}/dominant_id"])
    assert (out["footprint_pixels"] > 0).sum() >= 40 and len(set(out["dominant_id"].tolist())) >= 10


@pytest.mark.parametrize("name", ["ids_from_minus1", "ids_from_3"])
def test_sam_footprint_host_functions_vs_reference_golden(name):
    """opengaussian_b200.sam_footprints.fix_image / rgb_to_weight_map / most_common_id_weighted (the torch side of
    get_splat_id_and_weights) == the reference's own functions, on one-splat renders of the C oracle."""
    import helpers
    from opengaussian_b200 import sam_footprints as sf
    from oracle import raster as orc
    m, gold = _footprint_golden_module(), np.load(FGOLD)
    gs, cam, sam = m.inputs(name)
    g = helpers.np_inputs(gs)
    ocam = helpers.to_oracle_cam(cam)
    white = np.zeros((1, 16, 3), np.float32)
    white[:, 0, :] = 1.0
    assert abs(sf.WHITE_SH_COLOR - 0.7820948) < 1e-6
    for i in range(0, g["means3D"].shape[0], 3):
        sl = slice(i, i + 1)
        st = orc.forward(ocam, g["means3D"][sl], g["opacities"][sl], g["scales"][sl], g["rotations"][sl], shs=white,
                         bg=np.zeros(3, np.float32))
        img = sf.fix_image(torch.from_numpy(st.color[:3].copy()))
        assert img.dtype == torch.uint8 and tuple(img.shape) == (cam.image_height, cam.image_width, 3)
        w = sf.rgb_to_weight_map(img)
        assert tuple(w.shape) == (cam.image_height, cam.image_width, 1)
        assert int(img.max()) == gold[f"{name}/q_max"][i] and int((img != 0).any(2).sum()) == gold[f"{name}/footprint_pixels"][i]
        if gold[f"{name}/visible"][i]:
            assert float(w.max()) == 1.0
        assert sf.most_common_id_weighted(torch.from_numpy(sam), w) == gold[f"{name}/dominant_id"][i]
    # all-zero weights: index 0 of the shifted ids (the smallest id only if it is negative), one id everywhere: that id
    zero = torch.zeros(sam.shape[0], sam.shape[1], 1)
    lo = int(sam.min())
    assert sf.most_common_id_weighted(torch.from_numpy(sam), zero) == (lo if lo < 0 else 0)
    assert sf.most_common_id_weighted(torch.full((4, 5), 9), torch.zeros(4, 5, 1)) == 9


@pytest.mark.parametrize("mode", ["root", "leaf"])
def test_equalize_cluster_size_vs_reference_golden(mode):
    """Quantize_kMeans.equalize_cluster_size products (cluster_ids, cluster_len, max_cnt, excl_clusters,
    excl_cluster_ids) against what the reference's own method (scene/kmeans_quantize.py:89-144) produced on the same
    assignments (tests/golden/make_kmeans_golden.py): two clusters above the 10000-member threshold, an empty
    cluster, leaf mode with its k1*k2+1 rows.  Pure torch ops: runs on the CPU."""
    import importlib.util
    from opengaussian_b200.kmeans_quantize import Quantize_kMeans
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    spec = importlib.util.spec_from_file_location("mkg_eq", os.path.join(gdir, "make_kmeans_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    gold = np.load(os.path.join(gdir, "kmeans_golden.npz"))
    q = Quantize_kMeans(num_clusters=8, num_leaf_clusters=3, num_iters=1, dim=9)
    q.nn_index = torch.from_numpy(m.equalize_inputs(mode))
    q.equalize_cluster_size(mode=mode)
    assert (q.cls_ids if mode == "root" else q.leaf_cls_ids) is q.nn_index
    assert int(q.max_cnt) == int(gold[f"equalize_{mode}/max_cnt"]) and q.n_excl_cls == 2
    assert [int(e) for e in q.excl_clusters] == gold[f"equalize_{mode}/excl_clusters"].tolist()
    assert q.cluster_ids.dtype == torch.long and q.cluster_len.dtype == torch.long
    assert np.array_equal(q.cluster_ids.numpy(), gold[f"equalize_{mode}/cluster_ids"].astype(np.int64))
    assert np.array_equal(q.cluster_len.numpy(), gold[f"equalize_{mode}/cluster_len"])
    assert len(q.excl_cluster_ids) == 2
    for i, e in enumerate(q.excl_cluster_ids):
        assert np.array_equal(e.numpy(), gold[f"equalize_{mode}/excl_cluster_ids/{i}"].astype(np.int64))
