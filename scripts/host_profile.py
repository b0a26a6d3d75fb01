"""Developer probe: where does host time go in one fwd+bwd step (no cProfile distortion)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opengaussian_b200 import synth, _lib
import opengaussian_b200.rasterizer as R
from opengaussian_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer
gs, cams = synth.make_scene("lerf_1m_1080p", n_views=4)
dev = "cuda"
g = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in gs.items()}
P = g["means3D"].shape[0]
cam = cams[0].to(dev)
rs = GaussianRasterizationSettings(cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev), 1.0,
                                   cam.world_view_transform, cam.full_proj_transform, 3, cam.camera_center, False, False)
H, W = cam.image_height, cam.image_width
leaves = {k: g[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
m2 = torch.zeros(P, 3, device=dev, requires_grad=True)
G = torch.randn(5, H, W, device=dev)
rast = GaussianRasterizer(rs)
L = _lib.lib()
T = {"cfwd": 0.0, "cbwd": 0.0, "pyfwd": 0.0, "pybwd": 0.0, "zero": 0.0}
of, ob = L.ogs_raster_forward, L.ogs_raster_backward
class Wrap:
    def __init__(self, fn, key): self.fn, self.key = fn, key
    def __call__(self, *a):
        t0 = time.perf_counter(); r = self.fn(*a); T[self.key] += time.perf_counter() - t0; return r
class LW:
    def __getattr__(self, n):
        if n == "ogs_raster_forward": return Wrap(of, "cfwd")
        if n == "ogs_raster_backward": return Wrap(ob, "cbwd")
        return getattr(L, n)
R._lib.lib = lambda: LW()
_ob = R._RasterizeGaussians.backward
T["inner_bwd"] = 0.0
def _timed_bwd(ctx, *gs):
    t0 = time.perf_counter(); r = _ob(ctx, *gs); T["inner_bwd"] += time.perf_counter() - t0; return r
R._RasterizeGaussians.backward = staticmethod(_timed_bwd)
import gc
if os.environ.get("NOGC"): gc.disable()
def step():
    t0 = time.perf_counter()
    for t in list(leaves.values()) + [m2]:
        t.grad = None
    t1 = time.perf_counter()
    out = rast(means2D=m2, **leaves)
    t2 = time.perf_counter()
    torch.autograd.backward((out[0], out[2], out[3]), (G[0:3], G[3:4], G[4:5]))
    t3 = time.perf_counter()
    T["zero"] += t1 - t0; T["pyfwd"] += t2 - t1; T["pybwd"] += t3 - t2
for _ in range(5): step()
torch.cuda.synchronize()
for k in T: T[k] = 0.0
n = 30
t0 = time.perf_counter()
for _ in range(n): step()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
print("wall/step ms", wall / n * 1e3, {k: round(v / n * 1e3, 3) for k, v in T.items()})
