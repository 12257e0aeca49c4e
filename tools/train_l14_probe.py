"""Does the ViT-L/14@336 LoRA training step (BASELINE configs[4]) run?  One step at a small batch, r = 16."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import iic_b200
from importlib import import_module
clipc = import_module("ai-interior-image-classifier_b200.clip_compat"); lora = import_module("ai-interior-image-classifier_b200.lora")
B = int(os.environ.get("TRAIN_B", "32")); r = 16
dev = torch.device("cuda", 0)
vis = clipc.build_visual("ViT-L/14@336px", seed=0).to(dev)
for blk in vis.transformer.resblocks:
    blk.mlp.c_fc = lora.LoRALinear(blk.mlp.c_fc, rank=r, alpha=2 * r); blk.mlp.c_proj = lora.LoRALinear(blk.mlp.c_proj, rank=r, alpha=2 * r)
for n, p in vis.named_parameters():
    if n.endswith("lora_B"): p.data.normal_(0, 0.004)
images = torch.randint(0, 256, (B, 336, 336, 3), dtype=torch.uint8, device=dev)
text = torch.nn.functional.normalize(torch.randn(B, 768, device=dev), dim=-1)
tr = iic_b200.VisionLoRATrainer(vis, logit_scale=100.0)
losses = [tr.step(images, text) for _ in range(3)]
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): losses.append(tr.step(images, text))
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / 5 * 1e3
print(f"ViT-L/14@336 r=16 batch {B}: {ms:.1f} ms/step = {B / ms * 1e3:.0f} img/s; losses {[round(l, 4) for l in losses]}")
