import json, os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import iic_b200
from importlib import import_module
clipc = import_module("ai-interior-image-classifier_b200.clip_compat"); lora = import_module("ai-interior-image-classifier_b200.lora")
B = int(os.environ.get("TRAIN_B", "128"))
dev = torch.device("cuda", 0)
vis = clipc.build_visual("ViT-B/16", seed=0).to(dev)
for blk in vis.transformer.resblocks:
    blk.mlp.c_fc = lora.LoRALinear(blk.mlp.c_fc, rank=4, alpha=8); blk.mlp.c_proj = lora.LoRALinear(blk.mlp.c_proj, rank=4, alpha=8)
for n, p in vis.named_parameters():
    if n.endswith("lora_B"): p.data.normal_(0, 0.004)
images = torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8, device=dev)
text = torch.nn.functional.normalize(torch.randn(B, 512, device=dev), dim=-1)
tr = iic_b200.VisionLoRATrainer(vis, logit_scale=100.0)
for _ in range(3): tr.step(images, text)
torch.cuda.synchronize()
eng = tr.eng
eng.profile(True)
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
N = 5
for _ in range(N): tr.step(images, text)
e1.record(); torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / N * 1e3
prof = eng.profile_read()
print("gpu ms/step", e0.elapsed_time(e1) / N, "wall ms/step", wall)
print({k: (round(v["ms"] / N, 3), v["launches"] // N) for k, v in prof.items()})
# host-side cost of the pieces
for name, fn in (("_sync", tr._sync),):
    torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize(); print(name, (time.perf_counter() - t) * 1e3, "ms")
