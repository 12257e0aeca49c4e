"""BASELINE configs[3]: LoRA fine-tune step (fwd + bwd, LoRA-only grads, AdamW, NCCL all-reduce of the LoRA gradients),
batch 128 / GPU.  One process per GPU:
    python tools/train_bench.py                                     # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/train_bench.py
Prints one JSON line (rank 0): images/s of the training step, step time with and without all-reduce overlap, and a check
that every rank holds identical (averaged) gradients."""
import json, os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import iic_b200
from importlib import import_module
clipc = import_module("ai-interior-image-classifier_b200.clip_compat")
lora = import_module("ai-interior-image-classifier_b200.lora")


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    B = int(os.environ.get("TRAIN_B", "128")); steps = int(os.environ.get("STEPS", "10")); r = int(os.environ.get("RANK_LORA", "4"))
    mode = iic_b200._lib.operand_dtype_name()                # the one default operand dtype (IIC_OPERAND_DTYPE overrides)
    name = os.environ.get("MODEL", "ViT-B/16")              # or "ViT-L/14@336px" (BASELINE configs[4])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    vis = clipc.build_visual(name, seed=0).to(dev)
    R, E = vis.input_resolution, vis.output_dim
    vis.operand_dtype = mode
    for blk in vis.transformer.resblocks:
        blk.mlp.c_fc = lora.LoRALinear(blk.mlp.c_fc, rank=r, alpha=2 * r)
        blk.mlp.c_proj = lora.LoRALinear(blk.mlp.c_proj, rank=r, alpha=2 * r)
    torch.manual_seed(7)                       # identical LoRA init on every rank
    for n, p in vis.named_parameters():
        if n.endswith("lora_A"): p.data = (torch.randn(p.shape) * 0.02).to(dev)
        if n.endswith("lora_B"): p.data = (torch.randn(p.shape) * 0.004).to(dev)
    g = torch.Generator(device=dev).manual_seed(100 + rank)   # different data per rank
    images = torch.randint(0, 256, (B, R, R, 3), dtype=torch.uint8, device=dev, generator=g)
    text = torch.nn.functional.normalize(torch.randn(B, E, device=dev, generator=g), dim=-1)
    out = {}
    for overlap in (True, False):
        tr = iic_b200.VisionLoRATrainer(vis, logit_scale=100.0, overlap=overlap)
        for _ in range(3):
            tr.step(images, text)
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = tr.step(images, text)
        e1.record()
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["ms_per_step_overlap" if overlap else "ms_per_step_no_overlap"] = float(t)
        out["loss"] = loss
    # gradient consistency: after one forward_backward every rank must hold the same averaged gradients
    tr.forward_backward(images, text)
    flat = torch.cat([b for _, b in sorted(tr.buckets.items())])
    same = True
    if world > 1:
        ref = flat.clone(); dist.broadcast(ref, 0)
        same = bool(torch.equal(ref, flat))
        s = torch.tensor([1.0 if same else 0.0], device=dev); dist.all_reduce(s, op=dist.ReduceOp.MIN); same = bool(s.item() == 1.0)
    if rank == 0:
        ms = out["ms_per_step_overlap"]
        print(json.dumps({"metric": "LoRA fine-tune step images/s (fwd+bwd, LoRA-only grads, AdamW, NCCL all-reduce)", "value": world * B / (ms * 1e-3),
                          "unit": "images/s", "model": name, "n_gpus": world, "batch_per_gpu": B, "lora_rank": r, "operand_dtype": mode, **out,
                          "grad_bytes_allreduced_per_step": int(flat.numel() * 4), "ranks_hold_identical_grads": same,
                          "train_gflop_per_image": 72.1 if name == "ViT-B/16" else 810.2}), flush=True)
    if world > 1: dist.destroy_process_group()

main()
