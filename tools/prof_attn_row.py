"""One launch of each attention kernel at the BASELINE shape, for `ncu -k regex:attention` captures."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import iic_b200
eng = iic_b200.Engine(iic_b200.VIT_B_16, "cuda:0")
B, T, H = 1024, 197, 12
qkv = torch.randn(B * T, 3 * H * 64, device="cuda").to(eng.op_dtype)
impl = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(4):
    eng.op_attention(qkv, B, T, H, impl=impl)
torch.cuda.synchronize()
