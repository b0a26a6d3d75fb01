"""Multi-GPU check + timing of the in-switch gradient all-reduce (csrc/nvls.cu, dist.GradArena) against NCCL.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/nvls_check.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opengaussian_b200 import dist as ogd  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world}
    out["grid"], out["unroll"] = os.environ.get("OGS_NVLS_GRID"), os.environ.get("OGS_NVLS_UNROLL")
    for n in (59_000_320, 177_000_320)[:int(os.environ.get("NVLS_CASES", "2"))]:                      # 1 M and 3 M Gaussians x 59 floats (+ alignment)
        arena = ogd.GradArena(n)
        out["kind"] = arena.kind
        if arena.buf is None:
            out["error"] = getattr(arena, "error", "no multicast support")
            break
        g = torch.Generator(device=dev).manual_seed(rank)
        x = torch.randn(n, device=dev, generator=g)
        want = x.clone()
        dist.all_reduce(want)
        arena.buf[:n].copy_(x)
        arena.all_reduce(n)
        torch.cuda.synchronize()
        got = arena.buf[:n]
        out[f"max_abs_diff_{n}"] = float((got - want).abs().max())
        assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for label, fn in (("nvls", lambda: arena.all_reduce(n)), ("nvls_nobarrier", lambda: arena.all_reduce(n, barriers=False)),
                          ("nccl", lambda: dist.all_reduce(want))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            out[f"{label}_ms_{n * 4 // 1_000_000}MB"] = round(float(ms), 4)
        arena.close()
        del arena, x, want, got
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
