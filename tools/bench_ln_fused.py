"""out_proj / c_proj with the LayerNorm fused behind the residual epilogue vs GEMM + separate LayerNorm kernel (batch 1024)."""
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
M = int(os.environ.get("B", "1024")) * 197
for name, N, K in (("out_proj", 768, 768), ("c_proj", 768, 3072)):
    a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda"); x = torch.randn(M, N, device="cuda")
    gamma = torch.ones(N, device="cuda"); beta = torch.zeros(N, device="cuda"); la = torch.randn(N, 4, device="cuda") * 0.02
    t0 = timeit(lambda: eng.op_gemm(a, w, L.EPI_BIAS_RES_F32, bias=bias, residual=x, out=x))
    t1 = timeit(lambda: eng.op_layernorm(x, gamma, beta, lora_a_scaled=la))
    t3 = timeit(lambda: eng.op_gemm_res_ln(a, w, bias, x, gamma, beta, out=x))
    # the op wrapper allocates / frees its scratch per call: read the fused time as an upper bound (bench.py with IIC_FUSE_LN=1
    # measures it inside the encoder)
    print(f"{name:9s} gemm {t0:.3f} + layernorm {t1:.3f} = {t0 + t1:.3f} ms | fused (incl. scratch alloc) {t3:.3f} ms")
