"""Training-step A/B helper: the graph-replayed LoRA fine-tune step (batch 128, ViT-B/16, rank 4) timed as the median of
REPS blocks of STEPS steps, so that two builds (IIC_LIB=...) can be compared inside one GPU call at ~0.05 ms resolution."""
import os, sys, statistics, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import iic_b200
from importlib import import_module
clipc = import_module("ai-interior-image-classifier_b200.clip_compat"); lora = import_module("ai-interior-image-classifier_b200.lora")
B, steps, reps = int(os.environ.get("TRAIN_B", "128")), int(os.environ.get("STEPS", "20")), int(os.environ.get("REPS", "5"))
dev = torch.device("cuda", 0)
vis = clipc.build_visual("ViT-B/16", seed=0).to(dev)
for blk in vis.transformer.resblocks:
    blk.mlp.c_fc = lora.LoRALinear(blk.mlp.c_fc, rank=4, alpha=8); blk.mlp.c_proj = lora.LoRALinear(blk.mlp.c_proj, rank=4, alpha=8)
for n, p in vis.named_parameters():
    if n.endswith("lora_B"): p.data.normal_(0, 0.004)
images = torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8, device=dev)
text = torch.nn.functional.normalize(torch.randn(B, 512, device=dev), dim=-1)
tr = iic_b200.VisionLoRATrainer(vis, logit_scale=100.0, use_graph=os.environ.get("GRAPH", "1") == "1")
for _ in range(5): tr.step(images, text)
ts = []
for _ in range(reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): tr.step(images, text)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / steps)
print(os.environ.get("IIC_LIB", "default").split("/")[-1], "median %.3f ms  min %.3f  all %s" % (statistics.median(ts), min(ts), [round(t, 2) for t in ts]))
