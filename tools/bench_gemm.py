"""Micro-benchmark of the tcgen05 GEMM on the encoder shapes (B=1024 images -> M = 201728) vs cuBLAS (torch.matmul).
Writes gpurun_out/bench_gemm.json.  Each (ctas, shape) runs in this process; a trap aborts the script."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import iic_b200  # noqa: E402

L = iic_b200._lib


def timeit(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    B = int(os.environ.get("BENCH_B", "1024"))
    eng = iic_b200.Engine(iic_b200.VIT_B_16, "cuda:0")
    M = B * 197
    shapes = [("qkv", 2304, 768, L.EPI_BIAS_BF16), ("out", 768, 768, L.EPI_BIAS_RES_F32),
              ("fc", 3072, 768, L.EPI_BIAS_GELU_BF16), ("proj", 768, 3072, L.EPI_BIAS_RES_F32)]
    res = []
    for name, N, K, epi in shapes:
        a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
        bias = torch.randn(N, device="cuda")
        f32 = epi == L.EPI_BIAS_RES_F32
        out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
        flops = 2.0 * M * N * K
        t_cublas = timeit(lambda: torch.matmul(a, w.t()))
        row = {"name": name, "M": M, "N": N, "K": K, "cublas_ms": t_cublas, "cublas_tflops": flops / t_cublas / 1e9}
        for ctas in (1, 2):
            try:
                t = timeit(lambda: eng.op_gemm(a, w, epi, bias=bias, residual=out if f32 else None, out=out, ctas=ctas))
                row[f"cta{ctas}_ms"] = t
                row[f"cta{ctas}_tflops"] = flops / t / 1e9
            except Exception as ex:  # noqa: BLE001
                row[f"cta{ctas}_error"] = str(ex)[:200]
        print(row, flush=True)
        res.append(row)
        del a, w, out
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bench_gemm.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
