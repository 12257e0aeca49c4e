#!/usr/bin/env python
"""Benchmark of the hot path: images/sec of the LoRA-ViT image encoder + scoring head on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 1024] [--operand-dtype bf16|f16] [--model ViT-B/16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU implementation of the same path (oracle port)

One "step" = one pass of the whole hot path over one batch per GPU of synthetic uint8 224x224 images (BASELINE
configs[2]: batch 1024 / GPU, data parallel, no collective): preprocess kernel -> patch-embed GEMM -> 12 blocks
(LayerNorm, QKV GEMM, attention, out-proj GEMM, LayerNorm + LoRA-down, c_fc GEMM (+LoRA, QuickGELU), LoRA-down,
c_proj GEMM (+LoRA)) -> fused head over 437 labels in 6 groups.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec LoRA-ViT fwd at 1/2/4/8 B200 (bf16); tensor-pipe util vs peak"
UNIT = "images/s"
GROUPS = [40, 20, 12, 299, 36, 30]      # detector + styles, room_types, characteristics, materials, colors
SPLIT = [11, 0, 0, 0, 0, 0]

# algorithmic work per image, 2 FLOP per MAC
def flops_per_image(T, d, layers, patch_k, rank, embed):
    m = 4 * d
    patch = 2 * (T - 1) * patch_k * d
    per_layer_gemm = 2 * T * d * 3 * d + 2 * T * d * d + 2 * (2 * T * d * m)
    attn = 2 * (2 * T * T * d)
    lora = 2 * T * rank * (d + m) * 2
    total = patch + layers * (per_layer_gemm + attn + lora) + 2 * d * embed
    gemm_kernel = patch + layers * (per_layer_gemm + 2 * T * rank * (d + m))  # what the tcgen05 kernel itself computes
    return total, gemm_kernel


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"tflops": float(p.get("bf16_tflops_sustained", p.get("bf16_tflops"))), "hbm_gbs": float(p["hbm_gbs"]),
                "source": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md: ~1.4 PF sustained, 6.65 TB/s)"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        busy = [v for v in sm if v > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference(args, batch=16, budget_s=20.0):
    """The reference's CPU implementation of the path: the oracle restatement of CLIP + /root/reference/main.py's
    LoRALinear semantics (oracle/), fp32, all host threads, reference default batch 16 (main.py:371)."""
    import torch
    from oracle import clip_ref, ref_semantics as RS
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = clip_ref.ARCHS[args.model]
    torch.manual_seed(0)
    vis = clip_ref.VisionTransformer(cfg["image_resolution"], cfg["vision_patch_size"], cfg["vision_width"],
                                     cfg["vision_layers"], cfg["vision_width"] // 64, cfg["embed_dim"]).eval()
    RS.replace_linears_with_lora(vis, rank=args.lora_rank, alpha=2 * args.lora_rank)
    for n, p in vis.named_parameters():
        if n.endswith("lora_B"):
            p.data.normal_(0, 0.004)
    text = torch.nn.functional.normalize(torch.randn(sum(GROUPS), cfg["embed_dim"]), dim=-1)
    R = cfg["image_resolution"]
    x = torch.randn(batch, 3, R, R)

    def step():
        with torch.no_grad():
            f = vis(x)
            f = f / f.norm(dim=-1, keepdim=True)
            off = 0
            for n in GROUPS:
                (100.0 * f @ text[off:off + n].T).softmax(dim=-1).topk(min(5, n))
                off += n
    step()  # warm-up
    t0, n = time.perf_counter(), 0
    times = []
    while True:
        t1 = time.perf_counter(); step(); times.append(time.perf_counter() - t1); n += 1
        if (time.perf_counter() - t0 > budget_s and n >= 3) or n >= 50:
            break
    return {"value": batch / (sum(times) / len(times)), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} iterations of batch {batch} ({args.model}, fp32, LoRA r={args.lora_rank} on c_fc/c_proj, "
                      f"437-label head), oracle/clip_ref.py + oracle/ref_semantics.py"}, times


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    budget = min(25.0, 4.0 * steps)
    cb, times = cpu_reference(args, batch=16, budget_s=budget)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
            "warmup": 1, "ms_per_step": 1000.0 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"synthetic {args.model} LoRA-ViT inference, CPU fp32, batch 16 per step (bounded sample of "
                                   f"the batch-{args.batch}/GPU workload)", "model": args.model, "lora_rank": args.lora_rank},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=1024, help="images per GPU per step")
    ap.add_argument("--model", default="ViT-B/16", choices=["ViT-B/16", "ViT-L/14@336px"])
    ap.add_argument("--operand-dtype", default=os.environ.get("IIC_OPERAND_DTYPE", "bf16"), choices=["bf16", "f16"])
    ap.add_argument("--lora-rank", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import iic_b200
    from importlib import import_module
    clipc = import_module("ai-interior-image-classifier_b200.clip_compat")
    L = iic_b200._lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # stdout carries exactly ONE JSON line: anything libraries write to fd 1 meanwhile (NCCL prints its version banner
    # there) is sent to stderr, and fd 1 is restored just before the line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    K = max(args.steps, 1)

    # ---- synthetic workload ----
    torch.manual_seed(0)
    vis = clipc.build_visual(args.model, seed=0).to(dev)
    vis.operand_dtype = args.operand_dtype
    lora = import_module("ai-interior-image-classifier_b200.lora")
    for blk in vis.transformer.resblocks:       # LoRA on the vision MLPs: what main.py's wrap makes effective (F3/F4)
        blk.mlp.c_fc = lora.LoRALinear(blk.mlp.c_fc, rank=args.lora_rank, alpha=2 * args.lora_rank)
        blk.mlp.c_proj = lora.LoRALinear(blk.mlp.c_proj, rank=args.lora_rank, alpha=2 * args.lora_rank)
        for m in (blk.mlp.c_fc, blk.mlp.c_proj):
            m.lora.lora_B.data.normal_(0, 0.004)   # non-zero: the fused LoRA block cannot be skipped
    eng = vis.sync_engine()
    a = vis.arch
    E = a.embed_dim
    text = torch.nn.functional.normalize(torch.randn(sum(GROUPS), E, device=dev), dim=-1)
    eng.set_labels(text, GROUPS, SPLIT, topk=5, logit_scale=100.0)
    B, R = args.batch, a.image_size
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    images = torch.randint(0, 256, (B, R, R, 3), dtype=torch.uint8, device=dev, generator=g)
    host_images = torch.empty((B, R, R, 3), dtype=torch.uint8, pin_memory=True)
    host_images.copy_(images)
    staging = torch.empty_like(images)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return eng.classify_same_size(images, want_embedding=False)

    def step_e2e():
        return eng.classify_host_u8(host_images, staging)

    def run_e2e(n):
        """n batches through the streaming host-buffer API: every batch is copied H2D from pinned memory and its results are
        read back D2H inside the loop; the copy of batch i+1 overlaps the encode of batch i."""
        last = None
        for last in eng.classify_host_stream(host_images for _ in range(n)):
            pass
        return last

    # ---- device-resident timing: `value` ----
    for _ in range(W):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    eng.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        res = step_device()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    prof = eng.profile_read()
    eng.profile(False)
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end timing through the public call with host buffers: `e2e` ----
    step_e2e()
    run_e2e(2)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    tv, ti, ss = run_e2e(K)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    d2h = tv.numel() * 4 + ti.numel() * 4 + ss.numel() * 4

    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        peaks = measured_peaks()
        total_f, gemm_f = flops_per_image(a.tokens, a.width, a.layers, 3 * a.patch_size ** 2, args.lora_rank, E)
        gemm_ms = prof["gemm"]["ms"]
        gemm_launches = prof["gemm"]["launches"]
        ach = (gemm_f * B * K) / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("gemm_dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": world * B * K / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.operand_dtype == "bf16" else "f16", "data": "synthetic",
            "config": {"workload": f"synthetic {R}x{R} uint8 images -> preprocess + {args.model} LoRA-ViT encoder (r={args.lora_rank} on "
                                   f"mlp.c_fc/c_proj, non-zero B) + 437-label 6-group head, batch {B}/GPU, data parallel, no collective "
                                   "(BASELINE configs[2])",
                       "batch_per_gpu": B, "global_batch": world * B, "parallelism": f"dp{world}",
                       "operand_dtype": args.operand_dtype, "accumulate": "f32", "residual_stream": "f32",
                       "l2_policy": "per-step working set (~3.7 GB activations) >> 126 MB L2; no explicit flush",
                       "model_gflop_per_image": total_f / 1e9},
            "e2e": {"value": world * B * K / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(images.numel()),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / K,
                    "call": "Engine.classify_host_stream(pinned uint8 [B,R,R,3] batches) -> host top-k per batch (C ABI: iic_preprocess_same_size + "
                            "iic_classify; the H2D copy of batch i+1 overlaps the encode of batch i)"},
            "gpu_launches": int(sum(v["launches"] for k, v in prof.items() if not k.startswith("gemm_"))),
            "roofline": {"bound": "tensor", "kernel": "gemm_bf16_tn_kernel (tcgen05, all GEMM launches of the step)",
                         "achieved": ach, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": (ach / peaks["tflops"]) if ach else None,
                         "traffic": traffic, "peak_source": peaks["source"], "launches": gemm_launches,
                         "avg_launch_ms": gemm_ms / max(gemm_launches, 1),
                         "algorithmic_gflop_per_image": gemm_f / 1e9,
                         "model_tflops_whole_step": total_f * B * K / (ms_total * 1e-3) / 1e12 / 1.0,
                         "share_of_step": {k: v["ms"] / ms_total for k, v in prof.items() if not k.startswith("gemm_")},
                         "gemm_ms_per_launch_by_shape": {k[5:]: v["ms"] / max(v["launches"], 1) for k, v in prof.items()
                                                         if k.startswith("gemm_") and v["launches"]}},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"], _ = cpu_reference(args, batch=16, budget_s=15.0)
            except Exception as ex:  # noqa: BLE001
                line["cpu_baseline"] = {"error": str(ex)[:200]}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
