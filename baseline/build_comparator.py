"""Builds the GPU comparator baseline/upstream_structure.cu (a restatement of the upstream CUDA rasterizer's kernel
STRUCTURE, see the header of that file) into baseline/_build/libogs_upstream_structure.so.  The product's object
files are linked in for the shared preprocess kernels; the product library itself is untouched and never loads
this one.  Usage:  python baseline/build_comparator.py [--force]"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
SRC = os.path.join(HERE, "upstream_structure.cu")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libogs_upstream_structure.so")


def build(force: bool = False) -> str:
    from opengaussian_b200 import build as pb
    product = pb.build()
    objs = [os.path.join(pb.OBJ, u.replace(".cu", ".o")) for u in pb.UNITS]
    os.makedirs(OUT_DIR, exist_ok=True)
    obj = os.path.join(OUT_DIR, "upstream_structure.o")
    nvcc = pb._nvcc()
    deps = [SRC, os.path.join(pb.CSRC, "common.cuh"), os.path.join(ROOT, "include", "ogs_b200.h"), os.path.abspath(__file__)]
    if force or pb._stale(obj, deps):
        cmd = [nvcc] + pb.ARCH + pb.COMMON + ["-c", SRC, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj.replace(".o", ".ptxas.log"), "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if force or pb._stale(OUT, [obj, product] + objs):
        cmd = [nvcc] + pb.ARCH + ["-shared", "-o", OUT, obj] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
