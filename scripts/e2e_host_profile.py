"""Developer probe: cProfile of the host side of bench.py's end-to-end step (where does Python time go per view)."""
import cProfile
import os
import pstats
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

a = types.SimpleNamespace(workload="lerf_1m_1080p", views_per_step=4, streams=1)
cx = bench.Ctx()
wl = bench.RasterWorkload(cx, a.workload, 4, 1)
for i in range(6):
    wl.e2e_step(i)
wl.e2e_flush()
cx.torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(6, 26):
    wl.e2e_step(i)
wl.e2e_flush()
cx.torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
