"""train_lora.py's own training step (text tower through LoRA, image tower frozen under no_grad), batch 128 / GPU:
    python tools/train_text_bench.py                                     # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/train_text_bench.py
Each step = frozen vision encode of the image batch (engine inference path) + text forward / backward on the sequence engine
(causal attention, LoRA r=16 on the text MLPs: train_lora.py:62-100, 184) + clip_grad_norm_ + AdamW + NCCL all-reduce of the
LoRA gradients.  Prints one JSON line (rank 0)."""
import json, os, sys
os.environ.setdefault("IIC_ALLOW_RANDOM_INIT", "1")       # synthetic benchmark: seeded weights by design
os.environ.setdefault("IIC_ALLOW_STANDIN_TOKENIZER", "1")
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import iic_b200


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    B = int(os.environ.get("TRAIN_B", "128")); steps = int(os.environ.get("STEPS", "10")); r = int(os.environ.get("RANK_LORA", "16"))
    mode = os.environ.get("IIC_OPERAND_DTYPE", "bf16")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model, _ = iic_b200.load("ViT-B/16", device=dev, operand_dtype=mode)          # seeded random init (no checkpoints offline)
    iic_b200.replace_linears_with_lora(model, rank=r, alpha=2 * r)
    torch.manual_seed(7)
    for n, p in model.named_parameters():
        if n.startswith("transformer.") and n.endswith("lora_B"): p.data = (torch.randn(p.shape) * 0.004).to(dev)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    images = torch.randn(B, 3, 224, 224, device=dev, generator=g)
    tokens = torch.zeros(B, 77, dtype=torch.long, device=dev)
    n_tok = torch.randint(4, 76, (B,), device=dev, generator=g)
    body = torch.randint(1, 49405, (B, 77), device=dev, generator=g)
    pos = torch.arange(77, device=dev)[None]
    tokens = torch.where(pos < n_tok[:, None], body, tokens)
    tokens[:, 0] = 49406
    tokens[torch.arange(B, device=dev), n_tok] = 49407
    out = {}
    tr = iic_b200.TextLoRATrainer(model)
    for with_images in (True, False):
        feats = None
        if not with_images:
            with torch.no_grad():
                f = model.encode_image(images).float()
                feats = f / f.norm(dim=-1, keepdim=True)
        for _ in range(3):
            tr.step(images if with_images else feats, tokens)
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = tr.step(images if with_images else feats, tokens)
        e1.record()
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["ms_per_step" if with_images else "ms_per_step_text_only"] = float(t)
        out["loss"] = loss
    if rank == 0:
        ms = out["ms_per_step"]
        print(json.dumps({"metric": "train_lora.py step pairs/s (frozen vision encode + text-LoRA fwd+bwd, AdamW, NCCL all-reduce)",
                          "value": world * B / (ms * 1e-3), "unit": "pairs/s", "model": "ViT-B/16", "n_gpus": world, "batch_per_gpu": B,
                          "lora_rank": r, "operand_dtype": mode, **out}), flush=True)
    if world > 1: dist.destroy_process_group()

main()
