import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
import iic_b200
from torch.profiler import profile, ProfilerActivity
eng = iic_b200.Engine(iic_b200.VIT_B_16, "cuda:0")
B, T, H = 128, 197, 12
qkv = torch.randn(B * T, 3 * H * 64, device="cuda").to(eng.op_dtype)
do = torch.randn(B * T, H * 64, device="cuda").to(eng.op_dtype)
for _ in range(3): eng.op_attention_bwd(qkv, do, B, T, H)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): eng.op_attention_bwd(qkv, do, B, T, H)
    torch.cuda.synchronize()
for e in prof.key_averages():
    print(f"{e.key[:90]:90s} n={e.count} avg={e.device_time/1.0:.1f}us")
