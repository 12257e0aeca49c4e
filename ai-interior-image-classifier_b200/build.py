"""Builds csrc/*.cu into ONE in-tree shared library, `_lib/libiic_b200.so`, for sm_100a only.

    python ai-interior-image-classifier_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  cudart is linked statically and the driver API (cuTensorMapEncodeTiled) is
resolved at run time, so the library loads (and its symbols can be enumerated) on a CPU-only box.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(OUT_DIR, "libiic_b200.so")
SOURCES = ["gemm_sm100.cu", "rowwise.cu", "attention.cu", "attention_sm100.cu", "attention_row_sm100.cu", "attention_bwd.cu", "attention_bwd_sm100.cu", "attention_bwd_fused_sm100.cu", "train_ops.cu", "head.cu", "preprocess.cu", "jpeg.cu",
           "iic_api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
    "-Xcompiler", "-fPIC,-ffp-contract=off",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "iic.h"))
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    if not force and _newer(LIB, srcs + headers + [os.path.abspath(__file__)]):
        return LIB
    nvcc = _nvcc()

    def compile_one(src: str) -> str:
        obj = os.path.join(OUT_DIR, os.path.basename(src)[:-3] + ".o")
        if not force and _newer(obj, [src] + headers + [os.path.abspath(__file__)]):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("IIC_EXTRA_NVCC_FLAGS", "").split(), "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose and r.stderr:
            print(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-o", LIB, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
