"""CPU oracles (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this package.  ``opengaussian_b200`` never does.

* ``oracle.raster``  -- C restatement of the tile rasterizer (PARITY UNPINNED, see the
  header of ``raster_oracle.c``).
* ``oracle.kmeans``  -- C restatement of the codebook assign/accumulate + a Python mirror
  of ``Quantize_kMeans.cluster_assign`` control flow (pinned against the reference file).
* ``oracle.raster_torch`` -- independent fp64 autograd restatement used to validate the
  hand-derived backward of the C oracle.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile oracle/_build/libogs_oracle.so with gcc (needs only a C compiler)."""
    out = os.path.join(_HERE, "_build", "libogs_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("raster_oracle.c", "kmeans_oracle.c")]
    stale = force or not os.path.exists(out) or any(
        os.path.getmtime(s) > os.path.getmtime(out) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return out


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB
