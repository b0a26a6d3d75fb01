"""Driver of the GPU comparator (last row; baseline/upstream_structure.cu): loads the
upstream-STRUCTURE restatement of the reference CUDA rasterizer and times it against the product through the same
raw C-ABI calls.  Used by scripts/upstream_structure_bench.py, by bench.py's `gpu_comparator` leg and by
tests/test_comparator_gpu.py (parity of the comparator against the CPU oracle).  The comparator library is never
loaded by anything under opengaussian_b200/.

Two comparisons:
  * frame      -- one GaussianRasterizer forward + backward, SH degree 3, gradients to all inputs (the BASELINE metric);
  * stage1     -- OpenGaussian's instance-feature training step.  The reference (gaussian_renderer/__init__.py:104-163,
                  geometry detached at train.py:431-436) runs FOUR 3-channel forward passes (RGB, ins_feat[:, :3],
                  ins_feat[:, 3:6], silhouette) and TWO backward passes (the two feature passes; upstream's backward
                  always produces every gradient); the product runs ONE 9-channel forward and ONE colour-only backward.
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from opengaussian_b200 import _lib  # noqa: E402
from opengaussian_b200.rasterizer import GaussianRasterizationSettings, _Alloc, _fill_inputs  # noqa: E402

COMPARATOR = os.path.join(ROOT, "baseline", "_build", "libogs_upstream_structure.so")
NOTE = ("comparator = restatement of the upstream kernel STRUCTURE (one thread per pixel in 16x16 CTAs, one 64-bit-key "
        "library radix sort, 10 global atomics per contributing pixel-Gaussian pair); the reference's CUDA source is "
        "absent from the reference tree, so this is NOT the reference's code.  It shares the product's preprocess "
        "kernels (<1 % of the comparator's frame), which understates upstream's cost there.")


def load_comparator():
    if not os.path.exists(COMPARATOR):
        raise RuntimeError(f"{COMPARATOR} not found: run `python baseline/build_comparator.py`")
    L = C.CDLL(COMPARATOR)
    fwd_t, bwd_t = _lib.EXPORTS["ogs_raster_forward"], _lib.EXPORTS["ogs_raster_backward"]
    L.ups_raster_forward.restype, L.ups_raster_forward.argtypes = fwd_t
    L.ups_raster_backward.restype, L.ups_raster_backward.argtypes = bwd_t
    L.ogs_last_error.restype = C.c_char_p
    return L


def entry_points(which):
    """(forward, backward, last_error) of 'product' or 'upstream_structure'."""
    if which == "product":
        P_ = _lib.lib()
        return P_.ogs_raster_forward, P_.ogs_raster_backward, P_.ogs_last_error
    U = load_comparator()
    return U.ups_raster_forward, U.ups_raster_backward, U.ogs_last_error


class Pass:
    """One rasterizer pass (forward, optionally backward) through a (forward, backward, last_error) triple of C
    entry points on fixed inputs.  colours: 'sh' or a [P,3] tensor; extra: None or [P,F]; grads: 'all', 'colour'
    (product only: dL/dextra alone), or None (forward only)."""

    def __init__(self, fns, gs, cam, dev, colours="sh", extra=None, grads="all", bg=(0.1, 0.2, 0.3), seed=1):
        self.fwd, self.bwd, self.err = fns
        self.dev = dev
        t = lambda x: x.to(dev).float().contiguous()  # noqa: E731
        self.means3D, self.opac, self.scales, self.rots = t(gs["means3D"]), t(gs["opacities"]).reshape(-1), t(gs["scales"]), \
            t(gs["rotations"])
        self.shs = t(gs["shs"]) if isinstance(colours, str) else None
        self.colors = None if isinstance(colours, str) else t(colours)
        self.extra = None if extra is None else t(extra)
        self.F = 0 if extra is None else int(extra.shape[1])
        self.P = self.means3D.shape[0]
        self.H, self.W = cam.image_height, cam.image_width
        self.bg = torch.tensor(list(bg) + [0.0] * self.F, device=dev)
        self.rs = GaussianRasterizationSettings(self.H, self.W, cam.tanfovx, cam.tanfovy, self.bg[:3], 1.0,
                                                cam.world_view_transform.to(dev).contiguous(),
                                                cam.full_proj_transform.to(dev).contiguous(), 3,
                                                cam.camera_center.to(dev).contiguous(), False, False)
        self.grads_mode = grads
        z = lambda *s: torch.zeros(*s, device=dev)  # noqa: E731
        self.out = dict(color=z(3 + self.F, self.H, self.W), depth=z(self.H, self.W), alpha=z(self.H, self.W),
                        radii=torch.zeros(self.P, dtype=torch.int32, device=dev))
        g = torch.Generator(device=dev).manual_seed(seed)
        self.g_color = torch.randn(3 + self.F, self.H, self.W, device=dev, generator=g)
        self.g_depth = torch.randn(self.H, self.W, device=dev, generator=g)
        self.g_alpha = torch.randn(self.H, self.W, device=dev, generator=g)
        self.grads = {}
        if grads == "all":
            self.grads = dict(means3D=z(self.P, 3), means2D=z(self.P, 3), opacities=z(self.P), scales=z(self.P, 3),
                              rotations=z(self.P, 4))
            if self.shs is not None:
                self.grads["shs"] = z(self.P, 16, 3)
            else:
                self.grads["colors_precomp"] = z(self.P, 3)
            if self.F:
                self.grads["extra"] = z(self.P, self.F)
        elif grads == "colour":
            self.g_color[:3] = 0          # Stage 1: the loss sees the feature map only
            self.grads = dict(extra=z(self.P, self.F))
        if grads is not None:
            self.scratch = z(self.P * (3 + self.F + 7))
        self.num_rendered = 0

    def check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed (rc={rc}): {self.err().decode()}")

    def run(self):
        ri = _fill_inputs(self.rs, self.bg, self.means3D, self.opac, self.shs, self.colors, self.scales, self.rots, None,
                          self.extra, self.F)
        ro = _lib.RasterOutputs(*(self.out[k].data_ptr() for k in ("color", "depth", "alpha", "radii")))
        st = _lib.RasterState()
        alloc = _Alloc(self.dev)
        stream = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        self.check(self.fwd(C.byref(ri), C.byref(ro), alloc.fn, alloc.user, C.byref(st), stream), "forward")
        self.num_rendered = int(st.num_rendered)
        if self.grads_mode is not None:
            colour_only = self.grads_mode == "colour"
            gi = _lib.RasterGradsIn(self.g_color.data_ptr(), None if colour_only else self.g_depth.data_ptr(),
                                    None if colour_only else self.g_alpha.data_ptr())
            go = _lib.RasterGradsOut()
            for k, v in self.grads.items():
                setattr(go, "dL_d" + k, v.data_ptr())
            go.scratch = self.scratch.data_ptr()
            self.check(self.bwd(C.byref(ri), C.byref(st), C.byref(gi), C.byref(go), stream), "backward")
        self.bufs = alloc.bufs
        return self.num_rendered


def time_passes(passes, iters, warmup, dev):
    """Mean milliseconds for running every pass of `passes` once, CUDA events on the current stream."""
    for _ in range(warmup):
        for p in passes:
            p.run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(iters):
        for p in passes:
            p.run()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / iters


def frame_comparison(gs, cam, dev, iters=10, warmup=3):
    """fwd+bwd frame (SH degree 3, all gradients) through both libraries: {name: {ms_per_frame, frames_per_s}}."""
    res = {}
    for name in ("product", "upstream_structure"):
        p = Pass(entry_points(name), gs, cam, dev)
        ms = time_passes([p], iters, warmup, dev)
        res[name] = {"ms_per_frame": round(ms, 3), "frames_per_s": round(1e3 / ms, 1), "num_rendered": p.num_rendered}
        del p
        torch.cuda.empty_cache()
    res["speedup"] = round(res["upstream_structure"]["ms_per_frame"] / res["product"]["ms_per_frame"], 2)
    return res


def stage1_comparison(gs, cam, dev, iters=10, warmup=3):
    """Rasterizer work of one Stage-1 step: the reference's 4 forward + 2 backward 3-channel passes (upstream
    structure) against the product's one fused 9-channel forward + colour-only backward."""
    feat = (torch.nn.functional.normalize(gs["ins_feat"] * 2 - 1, dim=1) + 1) / 2     # gaussian_renderer/__init__.py:127
    ups = entry_points("upstream_structure")
    ref_passes = [Pass(ups, gs, cam, dev, grads=None),                                       # :104-112 RGB
                  Pass(ups, gs, cam, dev, colours=feat[:, :3], grads="all"),                # :129-138
                  Pass(ups, gs, cam, dev, colours=feat[:, 3:6], grads="all"),               # :141-151
                  Pass(ups, gs, cam, dev, grads=None)]                                       # :153-163 silhouette
    ms_ref = time_passes(ref_passes, iters, warmup, dev)
    del ref_passes
    torch.cuda.empty_cache()
    fused = Pass(entry_points("product"), gs, cam, dev, extra=feat, grads="colour")
    ms_fused = time_passes([fused], iters, warmup, dev)
    del fused
    torch.cuda.empty_cache()
    return {"upstream_structure_4fwd_2bwd_ms": round(ms_ref, 3), "product_fused_ms": round(ms_fused, 3),
            "speedup": round(ms_ref / ms_fused, 2)}
