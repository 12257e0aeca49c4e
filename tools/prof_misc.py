"""attention + layernorm at B=1024 shapes, for ncu (-k regex:attention_kernel|layernorm_kernel)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import iic_b200
B = int(os.environ.get("BENCH_B", "1024"))
eng = iic_b200.Engine(iic_b200.VIT_B_16, "cuda:0")
qkv = torch.randn(B * 197, 2304, device="cuda").to(torch.bfloat16)
x = torch.randn(B * 197, 768, device="cuda")
g = torch.randn(768, device="cuda"); b = torch.randn(768, device="cuda")
A = torch.randn(768, 4, device="cuda") * 0.04
for _ in range(2):
    eng.op_attention(qkv, B, 197, 12)
    eng.op_layernorm(x, g, b)
    eng.op_layernorm(x, g, b, lora_a_scaled=A)
torch.cuda.synchronize()
print("ok")
