"""Build libsift3d_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python -m 3dsift_b200.build        (or: python 3dsift_b200/build.py)

s3d_extract.cu is compiled with -fmad=false: the dense stages must round like the reference's
x86-64 build (separate FP32 multiply and add); s3d_match.cu keeps FMA contraction on (its
reference-exact arithmetic uses explicit __fmul_rn/__dadd_rn).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(HERE, "build")
LIB = os.path.join(OUT_DIR, "libsift3d_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden",
          "-I", os.path.join(ROOT, "include"), "-I", CSRC]

UNITS = [
    ("s3d_extract.cu", ["-fmad=false"]),
    ("s3d_slab.cu", []),
    ("s3d_match.cu", []),
    ("s3d_match_tc.cu", []),
    ("facade.cpp", []),
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers += [os.path.join(ROOT, "include", "sift3d_b200.h")]
    inc3 = os.path.join(ROOT, "include", "3dsift")
    if os.path.isdir(inc3):
        for d, _, fs in os.walk(inc3):
            headers += [os.path.join(d, f) for f in fs]
    objs = []
    for src, extra in UNITS:
        sp = os.path.join(CSRC, src)
        if not os.path.exists(sp):
            continue
        obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        if force or _stale(obj, [sp, __file__] + headers):
            cmd = [nvcc] + ARCH + COMMON + extra + os.environ.get("S3D_NVCC_EXTRA", "").split() + ["-x", "cu" if src.endswith(".cu") else "c++"]
            if verbose:
                cmd += ["-Xptxas", "-v"]
            cmd += ["-c", sp, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
    if force or _stale(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcuda", "-lz", "-ldl", "-lpthread"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
