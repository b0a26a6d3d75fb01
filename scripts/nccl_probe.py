"""Developer probe: all-reduce time of the gradient volume (236 MB fp32) on the box's GPUs."""
import os
import torch
import torch.distributed as dist

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sizes = {"shs 192 MB": 48_000_000, "flat 236 MB": 59_000_000, "small 12 MB": 3_000_000}
for name, n in sizes.items():
    x = torch.ones(n, device="cuda")
    for _ in range(5):
        dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    w = dist.get_world_size()
    if dist.get_rank() == 0:
        print(f"[{os.environ.get('TAG', '')}] {name}: {ms:.3f} ms  algbw {n * 4 / ms / 1e6:.0f} GB/s  busbw {n * 4 * 2 * (w - 1) / w / ms / 1e6:.0f} GB/s", flush=True)
dist.destroy_process_group()
