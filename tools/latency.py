"""Latency of the path at small batch (BASELINE configs[0]: one image): direct launches vs the CUDA-graph replay."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import iic_b200
from importlib import import_module
clipc = import_module("ai-interior-image-classifier_b200.clip_compat")
vis = clipc.build_visual("ViT-B/16", seed=0).cuda()
eng = vis.sync_engine()
text = torch.nn.functional.normalize(torch.randn(437, 512), dim=-1).cuda()
eng.set_labels(text, [40, 20, 12, 299, 36, 30], [11, 0, 0, 0, 0, 0], topk=5, logit_scale=100.0)
for B in (1, 4, 16):
    imgs = torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8).cuda()
    for graph in (False, True):
        for _ in range(5): eng.classify_same_size(imgs, use_graph=graph)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = 200
        for _ in range(n):
            r = eng.classify_same_size(imgs, use_graph=graph)
            r.topk_idx.cpu()                      # the caller reads the result: one sync per call
        dt = (time.perf_counter() - t0) / n * 1e3
        print(f"batch {B:2d} {'CUDA graph' if graph else 'direct    '}: {dt:.3f} ms per call (host wall clock, result read back)", flush=True)
