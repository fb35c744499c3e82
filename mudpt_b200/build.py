"""Build libmudpt_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m mudpt_b200.build [--bringup] [--force]

The shared library has no torch / libcuda link-time dependency (cudart is linked statically,
the driver entry point for TMA descriptors is resolved at run time), so it loads on the
CPU-only build box too -- every compute entry point then fails loudly (no CPU fallback).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
SOURCES = ["api.cu", "gemm.cu", "attention.cu", "attention_tc.cu", "rowops.cu", "head.cu", "prompt.cu", "augment.cu", "peer.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math", "-Xptxas", "-v"]
# --use_fast_math only affects the QuickGELU sigmoid / softmax exponentials inside epilogues whose
# results are rounded to bf16 anyway; the LayerNorm statistics use rsqrtf/div explicitly.


def lib_path(bringup: bool = False) -> str:
    return os.path.join(LIBDIR, "libmudpt_b200_bringup.so" if bringup else "libmudpt_b200.so")


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _digest(bringup: bool) -> str:
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ["../../include/mudpt_b200.h"]:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, "rb").read())
    h.update(repr(NVCC_FLAGS).encode() + (b"bringup" if bringup else b""))
    return h.hexdigest()


def build(bringup: bool = False, force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    out = lib_path(bringup)
    stamp = out + ".sha256"
    dig = _digest(bringup)
    if not force and os.path.exists(out) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return out
    objdir = os.path.join(LIBDIR, "obj_bringup" if bringup else "obj")
    os.makedirs(objdir, exist_ok=True)
    flags = list(NVCC_FLAGS) + (["-DMUDPT_BRINGUP"] if bringup else [])
    nvcc = _nvcc()

    def cc(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        r = subprocess.run([nvcc, *flags, "-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(cc, SOURCES))
    objs = [o for o, _ in results]
    with open(os.path.join(LIBDIR, "ptxas_bringup.log" if bringup else "ptxas.log"), "w") as f:
        for _, log in results:
            f.write(log)
    r = subprocess.run([nvcc, "-shared", "-o", out, *objs, "-cudart", "static", "-lpthread", "-ldl", "-lrt"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    return out


if __name__ == "__main__":
    force = "--force" in sys.argv
    p = build(bringup="--bringup" in sys.argv, force=force, verbose="-v" in sys.argv)
    print(p)
