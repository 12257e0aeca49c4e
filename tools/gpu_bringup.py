"""Bring-up runner for a GPU box: runs groups of GPU tests in separate processes (a trapped kernel poisons only
its own CUDA context), each under a timeout, and writes a summary to gpurun_out/bringup.json.

    python tools/gpu_bringup.py [group ...]
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
GROUPS = [
    ("gemm_1cta_basic", ["tests/test_kernels_gpu.py", "-k", "test_gemm_bias_bf16 and cta1 and shape0"]),
    ("gemm_2cta_basic", ["tests/test_kernels_gpu.py", "-k", "test_gemm_bias_bf16 and cta2 and shape0"]),
    ("gemm_1cta", ["tests/test_kernels_gpu.py", "-k", "test_gemm and cta1"]),
    ("gemm_2cta", ["tests/test_kernels_gpu.py", "-k", "test_gemm and cta2"]),
    ("layernorm", ["tests/test_kernels_gpu.py", "-k", "test_layernorm"]),
    ("attention", ["tests/test_kernels_gpu.py", "-k", "test_attention"]),
    ("head", ["tests/test_kernels_gpu.py", "-k", "test_head"]),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    want = set(sys.argv[1:])
    summary = {}
    for name, args in GROUPS:
        if want and name not in want:
            continue
        t0 = time.time()
        log = os.path.join(OUT, f"bringup_{name}.log")
        cmd = [sys.executable, "-m", "pytest", "-q", "-m", "gpu", "--no-header", "-p", "no:cacheprovider", *args]
        try:
            r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=420)
            rc, out = r.returncode, r.stdout + "\n" + r.stderr
        except subprocess.TimeoutExpired as e:
            rc, out = -999, (e.stdout or b"").decode(errors="replace") + "\nTIMEOUT"
        with open(log, "w") as f:
            f.write(out)
        tail = [l for l in out.strip().splitlines() if l.strip()][-3:]
        summary[name] = {"rc": rc, "secs": round(time.time() - t0, 1), "tail": tail}
        print(name, rc, tail[-1] if tail else "", flush=True)
        with open(os.path.join(OUT, "bringup.json"), "w") as f:
            json.dump(summary, f, indent=1)


if __name__ == "__main__":
    main()
