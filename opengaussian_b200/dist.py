"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL over NVLink on the B200 box,
gloo in the CPU tests).  The reference has no distributed code at all; the
hot path shards in exactly two places (section 8e):

* rasterizer: over CAMERA VIEWS -- every rank renders its own views against a full replica of the
  Gaussians; per-parameter gradients are summed with all-reduce (`allreduce_gradients`);
* k-means: over POINTS -- every rank owns a contiguous shard, centres are replicated, and one
  all-reduce of the packed [k*D sums | k counts] buffer is done per Lloyd iteration
  (`shard_kmeans` switches a Quantize_kMeans instance into that mode; ids stay sharded,
  `gather_ids` collects them when a caller such as save_kmeans needs the full vector).

No collective is invented where the path does not exchange data: a single frame is never split
across GPUs.
"""
from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def world(group=None):
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def split_views(views: Sequence, group=None) -> List:
    """Round-robin shard of a view list: rank r gets views r, r+R, r+2R, ..."""
    r, w = world(group)
    return list(views[r::w])


def shard_range(n: int, group=None):
    """Contiguous [lo, hi) shard of n items for this rank (k-means points)."""
    r, w = world(group)
    per = (n + w - 1) // w
    return min(n, r * per), min(n, (r + 1) * per)


def allreduce_gradients(params: Iterable[torch.Tensor], group=None, bucket_bytes: int = 256 << 20,
                        average: bool = False):
    """Sums .grad of every parameter across ranks, packing small tensors into flat buckets
    (bucket size chosen for launch latency, not link count: NVSwitch gives every peer full
    bandwidth).  Parameters whose .grad is None on this rank contribute zeros."""
    r, w = world(group)
    if w == 1:
        return
    params = [p for p in params if p.requires_grad]
    bucket, size = [], 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        if len(bucket) == 1 and bucket[0].grad is not None and bucket[0].grad.is_contiguous():
            dist.all_reduce(bucket[0].grad, group=group)
            if average:
                bucket[0].grad.div_(w)
        else:
            flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float()
                              for p in bucket])
            dist.all_reduce(flat, group=group)
            if average:
                flat.div_(w)
            off = 0
            for p in bucket:
                n = p.numel()
                g = flat[off:off + n].view_as(p).to(p.dtype)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
                off += n
        bucket, size = [], 0

    for p in params:
        nbytes = p.numel() * 4
        if nbytes >= bucket_bytes:
            flush()
            bucket = [p]
            flush()
            continue
        if size + nbytes > bucket_bytes:
            flush()
        bucket.append(p)
        size += nbytes
    flush()


def shard_kmeans(codebook, group=None):
    """Puts a Quantize_kMeans into sharded mode: its forward() then sees only this rank's points."""
    r, w = world(group)
    codebook.process_group = group
    codebook.distributed = w > 1
    return codebook


def gather_ids(local_ids: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather of the sharded id vector (shards may differ in length)."""
    r, w = world(group)
    if w == 1:
        return local_ids
    n = torch.tensor([local_ids.numel()], dtype=torch.int64, device=local_ids.device)
    sizes = [torch.zeros_like(n) for _ in range(w)]
    dist.all_gather(sizes, n, group=group)
    m = int(max(int(s) for s in sizes))
    pad = torch.zeros(m, dtype=local_ids.dtype, device=local_ids.device)
    pad[:local_ids.numel()] = local_ids
    outs = [torch.zeros_like(pad) for _ in range(w)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:int(s)] for o, s in zip(outs, sizes)])


def render_views_backward(render_loss, views: Sequence, params: Iterable[torch.Tensor], group=None,
                          already_split: bool = False) -> Optional[torch.Tensor]:
    """One view-parallel step: this rank renders ITS views (`render_loss(view) -> scalar loss`),
    backpropagates, and the parameter gradients of all ranks are summed.  Equivalent to one
    process rendering every view and summing the losses."""
    mine = views if already_split else split_views(views, group)
    total = None
    for v in mine:
        loss = render_loss(v)
        loss.backward()
        total = loss.detach() if total is None else total + loss.detach()
    params = list(params)
    allreduce_gradients(params, group)
    r, w = world(group)
    if w > 1:
        if total is None:
            total = torch.zeros((), device=params[0].device)
        dist.all_reduce(total, group=group)
    return total
