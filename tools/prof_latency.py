"""Two direct-launch passes at batch 1 (for `ncu --metrics gpu__time_duration.sum`: per-kernel durations of the latency path)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import iic_b200
from importlib import import_module
clipc = import_module("ai-interior-image-classifier_b200.clip_compat")
vis = clipc.build_visual("ViT-B/16", seed=0).cuda()
eng = vis.sync_engine()
text = torch.nn.functional.normalize(torch.randn(437, 512), dim=-1).cuda()
eng.set_labels(text, [40, 20, 12, 299, 36, 30], [11, 0, 0, 0, 0, 0], topk=5, logit_scale=100.0)
imgs = torch.randint(0, 256, (int(os.environ.get("B", "1")), 224, 224, 3), dtype=torch.uint8).cuda()
for _ in range(3):
    eng.classify_same_size(imgs, use_graph=False)
torch.cuda.synchronize()
