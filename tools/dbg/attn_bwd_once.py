import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
import iic_b200
eng = iic_b200.Engine(iic_b200.VIT_B_16, "cuda:0")
B, T, H = 128, 197, 12
qkv = torch.randn(B * T, 3 * H * 64, device="cuda").to(eng.op_dtype)
do = torch.randn(B * T, H * 64, device="cuda").to(eng.op_dtype)
for _ in range(3): eng.op_attention_bwd(qkv, do, B, T, H)
torch.cuda.synchronize()
