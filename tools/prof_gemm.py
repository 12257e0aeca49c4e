"""Launches each encoder GEMM shape (B=1024 -> M=201728) twice with the 2-CTA tcgen05 kernel; meant to be run under
   ncu --set full -k regex:gemm_bf16_tn -c 8   (launches 1,3,5,7 are the warm ones: qkv, out, fc, proj)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import iic_b200
L = iic_b200._lib
B = int(os.environ.get("BENCH_B", "1024"))
ctas = int(os.environ.get("CTAS", "2"))
eng = iic_b200.Engine(iic_b200.VIT_B_16, "cuda:0")
M = B * 197
for name, N, K, epi in [("qkv", 2304, 768, L.EPI_BIAS_BF16), ("out", 768, 768, L.EPI_BIAS_RES_F32),
                        ("fc", 3072, 768, L.EPI_BIAS_GELU_BF16), ("proj", 768, 3072, L.EPI_BIAS_RES_F32)]:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    f32 = epi == L.EPI_BIAS_RES_F32
    out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
    for _ in range(2):
        eng.op_gemm(a, w, epi, bias=bias, residual=out if f32 else None, out=out, ctas=ctas)
    torch.cuda.synchronize()
    del a, w, out
print("ok")
