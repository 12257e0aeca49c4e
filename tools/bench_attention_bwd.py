"""Attention backward alone at the training shape (batch 128, T = 197): tcgen05 kernel vs the mma.sync one (IIC_ATTN_BWD_IMPL=1)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import iic_b200
def timeit(fn, warmup=5, iters=30, reps=3):
    best = 1e9
    for _ in range(reps):
        for _ in range(warmup): fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters): fn()
        e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / iters)
    return best
eng = iic_b200.Engine(iic_b200.VIT_B_16, "cuda:0")
for B, T, H in ((128, 197, 12), (1024, 197, 12)):
    qkv = torch.randn(B * T, 3 * H * 64, device="cuda").bfloat16()
    do = torch.randn(B * T, H * 64, device="cuda").bfloat16()
    fwd = timeit(lambda: eng.op_attention(qkv, B, T, H))
    both = timeit(lambda: eng.op_attention_bwd(qkv, do, B, T, H))
    print(f"B={B} T={T}: forward {fwd:.3f} ms, forward+backward op {both:.3f} ms -> backward ~{both - fwd:.3f} ms  "
          f"(impl {os.environ.get('IIC_ATTN_BWD_IMPL', '0')})", flush=True)
