"""LoRA gradient reductions at the training step's shapes (batch 128: M = 25216 rows), timed alone with CUDA events through
the C ABI (iic_op_lora_bwd: dB + dP from one pass over Y;  iic_op_lora_outer: dA^T = act(Y)^T . dP), each call including its
second-stage reduction launches.  Prints GB/s of the algorithmic bytes (one read of Y) per call.
    python tools/bench_lora_grads.py            [IIC_LIB=... for an A/B against another build]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import iic_b200
from importlib import import_module
eng_mod = import_module("ai-interior-image-classifier_b200.engine")

B, T = int(os.environ.get("TRAIN_B", "128")), 197
M = B * T
eng = iic_b200.Engine(iic_b200.VIT_B_16, "cuda:0")
dt = eng.op_dtype
lib, h = eng.lib, eng.h
stream = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = {}
for N in (3072, 768):
    for rank in (4, 16):
        P = torch.zeros(M, 16, device="cuda", dtype=dt); P[:, :rank] = torch.randn(M, rank, device="cuda").to(dt)
        Bm = torch.zeros(16, N, device="cuda", dtype=dt); Bm[:rank] = (torch.randn(rank, N, device="cuda") * 0.05).to(dt)
        Y = torch.randn(M, N, device="cuda").to(dt)
        db = torch.zeros(rank, N, device="cuda"); dp = torch.zeros(M, 16, device="cuda", dtype=dt); da = torch.zeros(N, rank, device="cuda")
        nbytes = int(lib.iic_op_lora_scratch_bytes(N, M))
        scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        calls = {
            "bwd": lambda: lib.iic_op_lora_bwd(h, P.data_ptr(), 16, Y.data_ptr(), N, M, Bm.data_ptr(), rank, 1.0, db.data_ptr(), dp.data_ptr(),
                                                scratch.data_ptr(), nbytes, stream),
            "outer": lambda: lib.iic_op_lora_outer(h, P.data_ptr(), 16, Y.data_ptr(), N, M, 0, rank, 1.0, 1, da.data_ptr(), scratch.data_ptr(),
                                                   nbytes, stream),
            "outer_gelu": lambda: lib.iic_op_lora_outer(h, P.data_ptr(), 16, Y.data_ptr(), N, M, 1, rank, 1.0, 1, da.data_ptr(),
                                                        scratch.data_ptr(), nbytes, stream),
        }
        for name, fn in calls.items():
            ts = []
            for it in range(8):
                flush.zero_()                       # cold L2, like ncu's per-launch numbers; the step itself sees Y partly L2-resident
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(); rc = fn(); e.record(); torch.cuda.synchronize()
                assert rc == 0, (name, rc)
                ts.append(s.elapsed_time(e))
            t = sorted(ts[2:])[len(ts[2:]) // 2]
            res[f"{name}_N{N}_r{rank}"] = {"us": round(t * 1e3, 1), "gbs": round(M * N * 2 / t / 1e6, 0), "scratch_mb": round(nbytes / 2 ** 20, 1)}
            print(f"{name:11s} N={N:5d} rank={rank:2d}  {t * 1e3:7.1f} us  {M * N * 2 / t / 1e6:7.0f} GB/s", flush=True)
print(json.dumps(res))
