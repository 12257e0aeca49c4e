"""torch.profiler view of one training step (which kernels outside the engine's own classes take time, host gaps)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import iic_b200
from importlib import import_module
clipc = import_module("ai-interior-image-classifier_b200.clip_compat"); lora = import_module("ai-interior-image-classifier_b200.lora")
B = int(os.environ.get("TRAIN_B", "128")); r = int(os.environ.get("RANK_LORA", "4"))
dev = torch.device("cuda", 0)
vis = clipc.build_visual("ViT-B/16", seed=0).to(dev)
for blk in vis.transformer.resblocks:
    blk.mlp.c_fc = lora.LoRALinear(blk.mlp.c_fc, rank=r, alpha=2 * r); blk.mlp.c_proj = lora.LoRALinear(blk.mlp.c_proj, rank=r, alpha=2 * r)
for n, p in vis.named_parameters():
    if n.endswith("lora_B"): p.data.normal_(0, 0.004)
images = torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8, device=dev)
text = torch.nn.functional.normalize(torch.randn(B, 512, device=dev), dim=-1)
tr = iic_b200.VisionLoRATrainer(vis, logit_scale=100.0)
for _ in range(3): tr.step(images, text)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): tr.step(images, text)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=12, max_name_column_width=60))
