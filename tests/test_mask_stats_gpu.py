"""GPU parity of the mask-statistics kernels (csrc/mask_stats.cu through the C ABI) against the golden
vectors produced by the reference's own functions and against the CPU oracle at a larger size.
Tolerances: sums are fp32 with a different (atomic) order: 1e-5 relative; gradients 1e-4 relative."""
import importlib.util
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "mask_stats_golden.npz")


def _gm():
    spec = importlib.util.spec_from_file_location("mmg", os.path.join(os.path.dirname(GOLD), "make_mask_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _close(a, b, rtol, what):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert np.abs(a - b).max() <= rtol * np.abs(b).max() + 1e-8, (what, np.abs(a - b).max(), np.abs(b).max())


@pytest.mark.parametrize("name", ["stage1_6ch", "no_image_mask", "rgb_3ch"])
def test_stage1_loss_vs_reference_golden(name):
    from opengaussian_b200.mask_stats import cohesion_loss, mask_feature_mean, pair_mask_feature_mean, separation_loss
    m, gold = _gm(), np.load(GOLD)
    feat_np, masks_np, img_np = m.inputs(name)
    feat = torch.from_numpy(feat_np).cuda().requires_grad_(True)
    masks = torch.from_numpy(masks_np).cuda()
    img = None if img_np is None else torch.from_numpy(img_np).cuda().requires_grad_(True)
    mean = mask_feature_mean(feat, masks, image_mask=img)
    lc = cohesion_loss(feat, masks, mean)
    ls = separation_loss(mean, 1000)
    (ls + 0.1 * lc).backward()
    _close(mean.detach().cpu(), gold[f"{name}/mean"], 1e-5, "mean")
    _close(lc.detach().cpu(), gold[f"{name}/cohesion"], 1e-5, "cohesion")
    _close(ls.detach().cpu(), gold[f"{name}/separation"], 1e-5, "separation")
    _close(feat.grad.cpu(), gold[f"{name}/dfeat"], 1e-4, "dfeat")
    if img is not None:
        _close(img.grad.cpu(), gold[f"{name}/dimg"], 1e-4, "dimg")
    m2, var, cnt = mask_feature_mean(feat.detach(), masks, return_var=True)
    _close(m2.cpu(), gold[f"{name}/mean_noimg"], 1e-5, "mean_noimg")
    _close(var.cpu(), gold[f"{name}/var"], 1e-5, "var")
    assert np.array_equal(cnt.cpu().numpy(), gold[f"{name}/cnt"])
    pm = pair_mask_feature_mean(feat.detach().unsqueeze(0).repeat(3, 1, 1, 1), masks[:3])
    _close(pm.cpu(), gold[f"{name}/pair_mean"], 1e-5, "pair_mean")


@pytest.mark.parametrize("level", [0, 3])
@pytest.mark.parametrize("how", ["registered", "derived", "int64_onehot"])
def test_sam_partition_vs_reference_golden(level, how):
    """The reference's get_SAM_mask_and_feat -> mask_feature_mean -> cohesion/separation chain on a 4-level SAM id map
    (level 0: 12 masks, the mask-row walk; level 3: 25 masks, the id-map walk), with the id map registered by the
    drop-in get_SAM_mask_and_feat, derived on the device from bool masks, or from the reference's int64 one-hot."""
    from opengaussian_b200 import mask_stats as ms
    m, gold = _gm(), np.load(GOLD)
    sam, feat_np, img_np = m.sam_inputs()
    k = f"sam_l{level}"
    mask_id, mask_bool, invalid = ms.get_SAM_mask_and_feat(torch.from_numpy(sam).cuda(), level=level)
    shape = tuple(int(v) for v in gold[f"{k}/mask_bool_shape"])
    gold_bool = np.unpackbits(gold[f"{k}/mask_bool"])[: int(np.prod(shape))].reshape(shape).astype(bool)
    assert mask_bool.dtype == torch.bool and np.array_equal(mask_bool.cpu().numpy(), gold_bool)
    assert np.array_equal(mask_id.cpu().numpy(), gold[f"{k}/mask_id"])
    assert np.array_equal(invalid.cpu().numpy(), gold[f"{k}/invalid_pix"])
    if how == "registered":
        assert ms._last["src"] is mask_bool and ms._last["set"].ids is not None
    elif how == "derived":
        mask_bool = mask_bool.clone()                      # a new tensor: the id map is rebuilt by ogs_mask_id_map
    else:
        mask_bool = torch.nn.functional.one_hot(mask_id, shape[0] + 1).permute(2, 0, 1)[1:]   # reference :146-148,182
    feat = torch.from_numpy(feat_np).cuda().requires_grad_(True)
    img = torch.from_numpy(img_np).cuda().requires_grad_(True)
    mean = ms.mask_feature_mean(feat, mask_bool, image_mask=img)
    used = ms._last["set"]
    assert ms._last["src"] is mask_bool
    if shape[0] >= ms.ID_MAP_MIN_MASKS or how == "registered":
        assert used.ids is not None and int(used.overlap.item()) == 0
        assert np.array_equal(used.ids.cpu().numpy().reshape(shape[1:]), gold[f"{k}/mask_id"] - 1)
    lc = ms.cohesion_loss(feat, mask_bool, mean)
    assert ms._last["set"] is used                         # the second pass reuses the step's mask set
    ls = ms.separation_loss(mean, 1000)
    (ls + 0.1 * lc).backward()
    _close(mean.detach().cpu(), gold[f"{k}/mean"], 1e-5, "mean")
    _close(lc.detach().cpu(), gold[f"{k}/cohesion"], 1e-5, "cohesion")
    _close(ls.detach().cpu(), gold[f"{k}/separation"], 1e-5, "separation")
    _close(feat.grad.cpu(), gold[f"{k}/dfeat"], 1e-4, "dfeat")
    _close(img.grad.cpu(), gold[f"{k}/dimg"], 1e-4, "dimg")
    _, var, cnt = ms.mask_feature_mean(feat.detach(), mask_bool, return_var=True)
    _close(var.cpu(), gold[f"{k}/var"], 1e-5, "var")
    assert np.array_equal(cnt.cpu().numpy(), gold[f"{k}/cnt"])


def test_overlapping_masks_ignore_the_id_map():
    """24 masks of which two overlap: the device-side flag sends every pass down the mask-row walk (vs the CPU oracle);
    an in-place edit of the masks invalidates the cached mask set."""
    from opengaussian_b200 import mask_stats as ms
    from oracle import mask_stats as oms
    C, H, W, M = 6, 70, 93, 24
    g = torch.Generator().manual_seed(8)
    masks = _sam_like_masks(M, H, W, 9)
    masks[5] |= masks[3]
    masks[7, 10:30, 10:40] = True
    feat0 = torch.rand(C, H, W, generator=g)
    res = {}
    for dev in ("cpu", "cuda"):
        feat = feat0.clone().to(dev).requires_grad_(True)
        mk = masks.to(dev)
        fm, fc = (oms.mask_feature_mean, oms.cohesion_loss) if dev == "cpu" else (ms.mask_feature_mean, ms.cohesion_loss)
        mean = fm(feat, mk)
        (mean.pow(2).sum() + fc(feat, mk, mean)).backward()
        res[dev] = (mean.detach().cpu(), feat.grad.cpu())
        if dev == "cuda":
            assert int(ms._last["set"].overlap.item()) == 1
            first = ms._last["set"]
            mk[5] &= ~mk[3]
            mk[7, 10:30, 10:40] = False
            mk[7] &= ~(mk.sum(0) > 1)
            ms.mask_feature_mean(feat.detach(), mk)
            assert ms._last["set"] is not first
    _close(res["cuda"][0], res["cpu"][0], 2e-5, "mean")
    _close(res["cuda"][1], res["cpu"][1], 1e-4, "dfeat")


def test_id_map_walk_equals_mask_row_walk(monkeypatch):
    """Same partition masks through both walks (the id-map threshold forced out of reach for the second run)."""
    from opengaussian_b200 import mask_stats as ms
    C, H, W, M = 3, 211, 307, 40                         # H*W odd: ragged last lane, unaligned rows
    masks = _sam_like_masks(M, H, W, 5).cuda()
    masks[0] = False                                     # pixels in no mask at all
    g = torch.Generator().manual_seed(2)
    feat0 = torch.rand(C, H, W, generator=g).cuda()
    img0 = torch.rand(1, H, W, generator=g).cuda()
    out = []
    for thresh in (16, 10 ** 6):
        monkeypatch.setattr(ms, "ID_MAP_MIN_MASKS", thresh)
        ms._last.update(src=None, set=None)
        feat, img = feat0.clone().requires_grad_(True), img0.clone().requires_grad_(True)
        mean = ms.mask_feature_mean(feat, masks, image_mask=img)
        assert (ms._last["set"].ids is not None) == (thresh == 16)
        lc = ms.cohesion_loss(feat, masks, mean)
        (mean.pow(2).sum() + lc).backward()
        _, var, cnt = ms.mask_feature_mean(feat.detach(), masks, image_mask=img.detach(), return_var=True)
        out.append([t.detach().cpu() for t in (mean, lc, feat.grad, img.grad, var, cnt)])
    for a, b, what in zip(out[0], out[1], ("mean", "cohesion", "dfeat", "dimg", "var", "cnt")):
        _close(a, b, 2e-5, what)


def _sam_like_masks(M, H, W, seed):
    from opengaussian_b200 import synth
    return synth.sam_like_masks(M, H, W, seed)


def test_stage1_loss_vs_oracle_scannet_size():
    """BASELINE config 3 image size (1296x968), 120 SAM-like masks, 6 channels: CUDA vs CPU oracle."""
    from opengaussian_b200.mask_stats import cohesion_loss, mask_feature_mean
    from oracle import mask_stats as oms
    C, H, W, M = 6, 968, 1296, 120
    g = torch.Generator().manual_seed(3)
    feat0 = torch.rand(C, H, W, generator=g)
    img0 = (torch.rand(1, H, W, generator=g) > 0.1).float() * torch.rand(1, H, W, generator=g)
    masks = _sam_like_masks(M, H, W, 4)
    res = {}
    for dev in ("cpu", "cuda"):
        feat = feat0.detach().clone().to(dev).requires_grad_(True)
        img = img0.detach().clone().to(dev).requires_grad_(True)
        mk = masks.to(dev)
        if dev == "cpu":
            mean = oms.mask_feature_mean(feat, mk, image_mask=img)
            lc = oms.cohesion_loss(feat, mk, mean)
        else:
            mean = mask_feature_mean(feat, mk, image_mask=img)
            lc = cohesion_loss(feat, mk, mean)
        (mean.pow(2).sum() + lc).backward()
        res[dev] = (mean.detach().cpu(), lc.detach().cpu(), feat.grad.cpu(), img.grad.cpu())
    _close(res["cuda"][0], res["cpu"][0], 2e-5, "mean")
    _close(res["cuda"][1], res["cpu"][1], 2e-5, "cohesion")
    _close(res["cuda"][2], res["cpu"][2], 1e-4, "dfeat")
    _close(res["cuda"][3], res["cpu"][3], 1e-4, "dimg")


def test_empty_and_errors():
    from opengaussian_b200 import _lib
    from opengaussian_b200.mask_stats import cohesion_loss, mask_feature_mean
    feat = torch.rand(6, 16, 20, device="cuda", requires_grad=True)
    none = torch.zeros(0, 16, 20, dtype=torch.bool, device="cuda")
    assert mask_feature_mean(feat, none).shape == (0, 6)
    empty = torch.zeros(2, 16, 20, dtype=torch.bool, device="cuda")
    mean = mask_feature_mean(feat, empty)
    assert float(mean.detach().abs().max()) == 0.0                       # 0 / clamp(0, min=1)
    lc = cohesion_loss(feat, empty, mean)
    lc.backward()
    assert float(lc) == 0.0 and float(feat.grad.abs().max()) == 0.0
    with pytest.raises(_lib.OgsError):
        mask_feature_mean(torch.rand(5, 16, 20, device="cuda"), empty)      # unsupported channel count
    with pytest.raises(_lib.OgsError):
        mask_feature_mean(torch.rand(6, 16, 20), empty.cpu())                # no CPU path


@pytest.mark.parametrize("name", ["stage1_6ch", "no_image_mask", "rgb_3ch"])
def test_calculate_iou_vs_reference_golden(name):
    """csrc/mask_iou.cu through the C ABI == utils/opengs_utlis.py::calculate_iou, bit for bit (integer counts).
    stage1_6ch has H*W % 16 != 0 (byte path), the other two take the 16-byte path."""
    from opengaussian_b200.mask_stats import calculate_iou
    m, gold = _gm(), np.load(GOLD)
    m1, m2 = m.iou_inputs(name)
    a = torch.from_numpy(m1).cuda()
    b = torch.from_numpy(m2.astype(np.int32)).cuda()
    for base in (None, "former", "later"):
        iou = calculate_iou(a, b, base=base)
        assert tuple(iou.shape) == (m2.shape[0], m1.shape[0]) and iou.dtype == torch.float32
        assert np.array_equal(iou.cpu().numpy(), gold[f"{name}/iou_{base}"]), base


@pytest.mark.parametrize("H,W,n,m", [(968, 1296, 120, 10), (1080, 1920, 37, 19), (61, 67, 3, 1)])
def test_calculate_iou_vs_oracle_full_size(H, W, n, m):
    """Stage-3 association sizes (train.py:870): SAM-like masks against leaf silhouettes; exact counts."""
    from opengaussian_b200.mask_stats import calculate_iou, mask_pair_counts
    from oracle import mask_stats as oms
    m1 = _sam_like_masks(n, H, W, 5).cuda()
    m2 = torch.roll(_sam_like_masks(m, H, W, 6).cuda(), (7, -11), dims=(1, 2))
    inter, c1, c2 = mask_pair_counts(m1, m2)
    a = m1.view(n, -1).float()
    b = m2.view(m, -1).float()
    assert torch.equal(inter, (b @ a.t()).round().to(torch.int32))       # fp32 GEMM of 0/1: exact below 2^24
    assert torch.equal(c1, m1.view(n, -1).sum(1).to(torch.int32)) and torch.equal(c2, m2.view(m, -1).sum(1).to(torch.int32))
    for base in (None, "former", "later"):
        ref = oms.calculate_iou(m1.cpu(), m2.cpu(), base=base)
        assert torch.equal(calculate_iou(m1, m2, base=base).cpu(), ref), base
    # symmetry: IoU(a, b) == IoU(b, a)^T, and a set against itself has a unit diagonal (where non-empty)
    assert torch.equal(calculate_iou(m1, m2), calculate_iou(m2, m1).t())
    self_iou = calculate_iou(m1, m1).diagonal()
    assert torch.all((self_iou == 0) | ((self_iou - 1).abs() < 1e-6))


def test_calculate_iou_edge_cases():
    from opengaussian_b200 import _lib
    from opengaussian_b200.mask_stats import calculate_iou
    z = torch.zeros(0, 8, 8, dtype=torch.bool, device="cuda")
    o = torch.ones(2, 8, 8, dtype=torch.bool, device="cuda")
    assert tuple(calculate_iou(z, o).shape) == (2, 0) and tuple(calculate_iou(o, z).shape) == (0, 2)
    assert torch.equal(calculate_iou(o, torch.zeros_like(o)), torch.zeros(2, 2, device="cuda"))   # 0 / 1e-6
    v = o[:, ::2]                                                  # non-contiguous view
    assert torch.allclose(calculate_iou(v, v), torch.ones(2, 2, device="cuda"))
    with pytest.raises(_lib.OgsError):
        calculate_iou(o.cpu(), o.cpu())


@pytest.mark.parametrize("M,H,W", [(1, 16, 16), (7, 48, 64), (120, 968, 1296), (254, 64, 80), (255, 64, 80), (300, 32, 48),
                                   (12, 33, 35)])
def test_sam_mask_expansion_equals_one_hot(M, H, W):
    """get_SAM_mask_and_feat's [num_mask,H,W] bool masks (reference utils/opengs_utlis.py:146-148: one_hot of the id map,
    mask 0 dropped) from the byte-parallel expansion kernel (M <= 254, H*W % 16 == 0) and from the per-mask kernel."""
    from opengaussian_b200 import mask_stats as ms
    g = torch.Generator().manual_seed(M * 131 + H)
    level_ids = torch.randint(-1, M, (H, W), generator=g)             # the SAM file's ids: -1 = no mask, 0..M-1 = masks
    level_ids[0, :4] = torch.tensor([0, min(1, M - 1), -1, M - 1])    # a 0x01 byte above a 0x00 byte, "no mask", the last id
    sam = torch.stack([level_ids, level_ids, level_ids, level_ids]).cuda()
    mask_id, mask_bool, invalid = ms.get_SAM_mask_and_feat(sam, level=0, num_mask=M)
    want_id = level_ids.cuda() + 1                                    # reference :142: 0 = invalid, 1..M
    want = torch.nn.functional.one_hot(want_id, M + 1).permute(2, 0, 1)[1:].bool()
    assert mask_bool.shape == (M, H, W) and mask_bool.dtype == torch.bool
    assert torch.equal(mask_bool, want)
    assert torch.equal(mask_bool.view(torch.uint8), want.view(torch.uint8))      # bytes are exactly 0 / 1
    assert torch.equal(invalid, want_id == 0) and torch.equal(mask_id, want_id)
