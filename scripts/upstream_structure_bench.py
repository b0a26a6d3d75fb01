"""GPU comparator run (last row): the upstream-STRUCTURE restatement of the reference CUDA
rasterizer (baseline/upstream_structure.cu -- one thread per pixel, 64-bit key library sort, per-pair global
atomics; NOT the reference's code, whose source is absent) against the product, through the same raw C-ABI calls:
  1. parity on a small scene: images 1e-5 absolute (expf vs ex2.approx: 2e-5 allowed), gradients 1e-3 relative;
  2. fwd+bwd frames/s of both on the BASELINE workload (1 M Gaussians, 1920x1080, SH degree 3, all gradients).
Prints one JSON line.  Needs a GPU; run `python baseline/build_comparator.py` first (done by gpu_evidence.sh)."""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opengaussian_b200 import _lib, synth  # noqa: E402
from opengaussian_b200.rasterizer import GaussianRasterizationSettings, _Alloc, _fill_inputs  # noqa: E402

COMPARATOR = os.path.join(ROOT, "baseline", "_build", "libogs_upstream_structure.so")


def load_comparator():
    L = C.CDLL(COMPARATOR)
    fwd_t, bwd_t = _lib.EXPORTS["ogs_raster_forward"], _lib.EXPORTS["ogs_raster_backward"]
    L.ups_raster_forward.restype, L.ups_raster_forward.argtypes = fwd_t
    L.ups_raster_backward.restype, L.ups_raster_backward.argtypes = bwd_t
    L.ogs_last_error.restype = C.c_char_p
    return L


class Frame:
    """One forward + backward through a (forward, backward, last_error) triple of C entry points."""

    def __init__(self, fwd, bwd, err, gs, cam, dev):
        self.fwd, self.bwd, self.err, self.dev = fwd, bwd, err, dev
        t = lambda k: gs[k].to(dev).float().contiguous()  # noqa: E731
        self.means3D, self.opac, self.scales, self.rots, self.shs = t("means3D"), t("opacities").reshape(-1), t("scales"), \
            t("rotations"), t("shs")
        self.P = self.means3D.shape[0]
        self.H, self.W = cam.image_height, cam.image_width
        self.bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
        self.rs = GaussianRasterizationSettings(self.H, self.W, cam.tanfovx, cam.tanfovy, self.bg, 1.0,
                                                cam.world_view_transform.to(dev).contiguous(),
                                                cam.full_proj_transform.to(dev).contiguous(), 3,
                                                cam.camera_center.to(dev).contiguous(), False, False)
        g = torch.Generator(device=dev).manual_seed(1)
        self.g_color = torch.randn(3, self.H, self.W, device=dev, generator=g)
        self.g_depth = torch.randn(self.H, self.W, device=dev, generator=g)
        self.g_alpha = torch.randn(self.H, self.W, device=dev, generator=g)
        z = lambda *s: torch.zeros(*s, device=dev)  # noqa: E731
        self.out = dict(color=z(3, self.H, self.W), depth=z(self.H, self.W), alpha=z(self.H, self.W),
                        radii=torch.zeros(self.P, dtype=torch.int32, device=dev))
        self.grads = dict(means3D=z(self.P, 3), means2D=z(self.P, 3), opacities=z(self.P), shs=z(self.P, 16, 3),
                          scales=z(self.P, 3), rotations=z(self.P, 4))
        self.scratch = z(self.P * 10)

    def check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed (rc={rc}): {self.err().decode()}")

    def step(self):
        ri = _fill_inputs(self.rs, self.bg, self.means3D, self.opac, self.shs, None, self.scales, self.rots, None, None, 0)
        ro = _lib.RasterOutputs(*(self.out[k].data_ptr() for k in ("color", "depth", "alpha", "radii")))
        st = _lib.RasterState()
        alloc = _Alloc(self.dev)
        stream = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        self.check(self.fwd(C.byref(ri), C.byref(ro), alloc.fn, None, C.byref(st), stream), "forward")
        gi = _lib.RasterGradsIn(self.g_color.data_ptr(), self.g_depth.data_ptr(), self.g_alpha.data_ptr())
        go = _lib.RasterGradsOut()
        go.dL_dmeans3D, go.dL_dmeans2D, go.dL_dopacities = (self.grads[k].data_ptr() for k in ("means3D", "means2D", "opacities"))
        go.dL_dshs, go.dL_dscales, go.dL_drotations = (self.grads[k].data_ptr() for k in ("shs", "scales", "rotations"))
        go.scratch = self.scratch.data_ptr()
        self.check(self.bwd(C.byref(ri), C.byref(st), C.byref(gi), C.byref(go), stream), "backward")
        self.bufs = alloc.bufs
        return int(st.num_rendered)

    def timed(self, n, warm):
        for _ in range(warm):
            self.step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            self.step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="lerf_1m_1080p")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    P_ = _lib.lib()
    U = load_comparator()
    prod = (P_.ogs_raster_forward, P_.ogs_raster_backward, P_.ogs_last_error)
    comp = (U.ups_raster_forward, U.ups_raster_backward, U.ogs_last_error)

    # 1. parity on the plumbing scene
    gs, cams = synth.make_scene("plumbing_10k_256", n_views=2)
    fa, fb = Frame(*prod, gs, cams[0], dev), Frame(*comp, gs, cams[0], dev)
    na, nb = fa.step(), fb.step()
    torch.cuda.synchronize()
    parity = {"num_rendered_equal": na == nb, "radii_equal": bool(torch.equal(fa.out["radii"], fb.out["radii"]))}
    for k in ("color", "depth", "alpha"):
        parity[f"{k}_max_abs"] = float((fa.out[k] - fb.out[k]).abs().max())
    for k in fa.grads:
        parity[f"d{k}_rel"] = float((fa.grads[k] - fb.grads[k]).abs().max() / (fa.grads[k].abs().max() + 1e-20))
    ok = parity["num_rendered_equal"] and parity["radii_equal"] and \
        all(parity[f"{k}_max_abs"] <= 2e-5 * max(1.0, float(fa.out[k].abs().max())) for k in ("color", "depth", "alpha")) and \
        all(parity[f"d{k}_rel"] <= 1e-3 for k in fa.grads)
    del fa, fb

    # 2. timing on the BASELINE workload
    gs, cams = synth.make_scene(a.scene, n_views=2)
    res = {}
    for name, fns in (("product", prod), ("upstream_structure", comp)):
        f = Frame(*fns, gs, cams[0], dev)
        ms = f.timed(a.iters, a.warmup)
        res[name] = {"ms_per_frame": round(ms, 3), "frames_per_s": round(1e3 / ms, 1)}
        del f
        torch.cuda.empty_cache()
    print(json.dumps({"probe": "upstream_structure_comparator", "scene": a.scene, "parity_ok": bool(ok), "parity": parity,
                      **res, "speedup": round(res["upstream_structure"]["ms_per_frame"] / res["product"]["ms_per_frame"], 2),
                      "note": "comparator = restatement of the upstream kernel STRUCTURE (per-pixel threads, one 64-bit "
                              "library sort, per-pair atomics) sharing the product's preprocess kernels; not the reference's code"}))


if __name__ == "__main__":
    main()
