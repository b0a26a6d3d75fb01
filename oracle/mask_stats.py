"""CPU restatement (torch, float64 accumulation) of the reference's per-mask feature statistics and
Stage-1 losses -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows utils/opengs_utlis.py::mask_feature_mean (:240-283) and train.py::cohesion_loss (:102-121) /
separation_loss (:123-155) without their [M,C,H,W] expansion, and utils/opengs_utlis.py::calculate_iou
(:90-123) as integer pixel counts.  Pinned against the reference functions
themselves: tests/golden/make_mask_golden.py runs them (torch CPU) and commits outputs and autograd
gradients; tests/test_oracle_cpu.py checks this restatement against those fixtures.
"""
import torch


def mask_feature_mean(feat_map, gt_masks, image_mask=None, return_var=False):
    """feat_map [C,H,W], gt_masks [M,H,W] bool, image_mask [1,H,W] float or None (reference :240-283)."""
    C, H, W = feat_map.shape
    f = feat_map.double().reshape(C, -1)                       # [C, HW]
    m = gt_masks.reshape(gt_masks.shape[0], -1).double()       # [M, HW]
    w = m if image_mask is None else m * image_mask.double().reshape(1, -1)   # :253-255
    sums = w @ f.t()                                           # [M, C]  sum_p feat * mask * image_mask
    counts = w.sum(1).clamp(min=1)                             # :263
    mean = sums / counts[:, None]                              # :267
    if not return_var:
        return mean.to(feat_map.dtype)
    fw = f if image_mask is None else f * image_mask.double().reshape(1, -1)
    # :270-276  sum over the mask of (masked_feats - mean)^2 / counts, then the mean over channels
    sq = torch.stack([((fw - mean[i][:, None]) ** 2 * m[i][None, :]).sum(1) for i in range(m.shape[0])]) \
        if m.shape[0] else f.new_zeros(0, C)
    variance = (sq / counts[:, None]).mean(1)
    return mean.to(feat_map.dtype), variance.to(feat_map.dtype), counts.to(feat_map.dtype)


def cohesion_loss(feat_map, gt_mask, feat_mean_stack):
    """Reference train.py:102-121."""
    C = feat_map.shape[0]
    f = feat_map.double().reshape(C, -1)
    m = gt_mask.reshape(gt_mask.shape[0], -1).double()
    mu = feat_mean_stack.double()
    per_mask = []
    for i in range(m.shape[0]):
        dist = (f * m[i][None, :] - mu[i][:, None]).norm(p=2, dim=0)      # :113-114
        per_mask.append((dist * m[i]).sum() / m[i].sum().clamp(min=1))     # :117-118
    return torch.stack(per_mask).mean().to(feat_map.dtype)                # :120


def separation_loss(feat_mean_stack, iteration):
    """Reference train.py:123-155."""
    N = feat_mean_stack.shape[0]
    diff_squared = (feat_mean_stack.unsqueeze(1) - feat_mean_stack.unsqueeze(0)).pow(2).sum(2)
    inverse_distance = 1.0 / (diff_squared + 1)
    inverse_distance = inverse_distance.masked_fill(torch.eye(N, device=feat_mean_stack.device).bool(), 0)
    sorted_indices = inverse_distance.argsort().argsort()
    loss_weight = (sorted_indices.float() / (N - 1)) * (1.0 - 0.1) + 0.1
    if iteration > 35_000:
        loss_weight[loss_weight < 0.9] = 0.1
    return (inverse_distance * loss_weight).sum() / (N * (N - 1))


def calculate_iou(masks1, masks2, base=None):
    """Reference utils/opengs_utlis.py:90-123: masks1 [n,H,W], masks2 [m,H,W] -> IoU [m,n].  Pixel counts are
    integers (exact in the reference's float32 sums below 2^24), the division is done in float32 like :120."""
    a = (masks1 != 0).reshape(masks1.shape[0], -1).to(torch.int64)          # [n, HW]
    b = (masks2 != 0).reshape(masks2.shape[0], -1).to(torch.int64)          # [m, HW]
    inter = b @ a.t()                                                       # :110
    ca, cb = a.sum(1), b.sum(1)
    if base == "former":
        union = ca[None, :].float() + 1e-6                                  # :114
    elif base == "later":
        union = cb[:, None].float() + 1e-6                                  # :116
    else:
        union = (ca[None, :] + cb[:, None] - inter).float() + 1e-6          # :118  |a or b| = |a| + |b| - |a and b|
    return inter.float() / union
