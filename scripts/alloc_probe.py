import os, time, torch
print("ALLOC_CONF", os.environ.get("PYTORCH_CUDA_ALLOC_CONF"), os.environ.get("PYTORCH_ALLOC_CONF"), torch.cuda.get_allocator_backend())
dev="cuda"
def t(n, reps=200):
    xs=[torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(3)]
    del xs
    torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(reps):
        x=torch.empty(n, dtype=torch.uint8, device=dev); del x
    return (time.perf_counter()-t0)/reps*1e6
for n in [1024, 1<<20, 12<<20, 33<<20, 192<<20]:
    print(n, "us/empty", round(t(n),1))
import ctypes
