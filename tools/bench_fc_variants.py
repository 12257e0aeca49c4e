import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import iic_b200
L = iic_b200._lib
def timeit(fn, warmup=3, iters=10):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters
eng = iic_b200.Engine(iic_b200.VIT_B_16, "cuda:0")
M, N, K = 1024 * 197, 3072, 768
a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
bias = torch.randn(N, device="cuda"); out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
p = torch.randn(M, 16, device="cuda").bfloat16(); bt = torch.randn(N, 16, device="cuda").bfloat16()
da = torch.randn(N, 4, device="cuda"); dpart = torch.zeros(24, M, 4, device="cuda")
for name, kw in [("bias only (epi 0)", dict(epilogue=L.EPI_BIAS_BF16)), ("gelu", dict(epilogue=L.EPI_BIAS_GELU_BF16)),
                 ("gelu+lora", dict(epilogue=L.EPI_BIAS_GELU_BF16, lora_p=p, lora_bt=bt, r_pad=16)),
                 ("gelu+down", dict(epilogue=L.EPI_BIAS_GELU_BF16, down_a=da, down_part=dpart)),
                 ("gelu+lora+down", dict(epilogue=L.EPI_BIAS_GELU_BF16, lora_p=p, lora_bt=bt, r_pad=16, down_a=da, down_part=dpart)),
                 ("erf gelu", dict(epilogue=L.EPI_GELU_ERF_BF16))]:
    t = timeit(lambda: eng.op_gemm(a, w, bias=bias, out=out, ctas=2, **kw))
    print(f"{name:22s} {t:.3f} ms  {2.0 * M * N * K / t / 1e9:.0f} TFLOP/s")
