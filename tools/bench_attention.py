"""Attention kernels alone: mma.sync vs tcgen05, ViT-B/16@224 (T=197) and ViT-L/14@336 (T=577) shapes."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import iic_b200
def timeit(fn, warmup=5, iters=30, reps=3):
    return min(_timeit(fn, warmup, iters) for _ in range(reps))
def _timeit(fn, warmup, iters):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters
eng = iic_b200.Engine(iic_b200.VIT_B_16, "cuda:0")
impls = [(1, "mma.sync"), (2, "tcgen05-blocks"), (3, "tcgen05-row")]
if len(sys.argv) > 1: impls = [i for i in impls if str(i[0]) in sys.argv[1:]]
for B, T, H in ((1024, 197, 12), (256, 577, 16)):
    qkv = torch.randn(B * T, 3 * H * 64, device="cuda").to(eng.op_dtype)
    for impl, name in impls:
        try:
            t = timeit(lambda: eng.op_attention(qkv, B, T, H, impl=impl))
            print(f"B={B} T={T} H={H} {name:10s} {t:.3f} ms   {4.0 * T * T * 64 * B * H / t / 1e9:.0f} TFLOP/s (unpadded)", flush=True)
        except Exception as ex:
            print(name, "FAILED", str(ex)[:200])
