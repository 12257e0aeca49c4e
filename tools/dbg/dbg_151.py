import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["IIC_ALLOW_RANDOM_INIT"] = "1"; os.environ["IIC_ALLOW_STANDIN_TOKENIZER"] = "1"
import numpy as np, torch
import iic_b200
from _common import golden_npz, golden_json, oracle_state_dict, label_layout
model, pre = iic_b200.load("ViT-B/16", device="cuda", state_dict=oracle_state_dict())
lab, sizes, split = label_layout()
text = torch.from_numpy(golden_npz("text_features.npz")["text"]).cuda()
crops = torch.from_numpy(golden_npz("crops_u8.npz")["crops"]).cuda()
ref = golden_npz("ref_shipped.npz")
ref_logits = torch.from_numpy(ref["logits"]).cuda()
eng = model.visual.sync_engine(use_lora=False)
eng.set_labels(text, sizes, split, topk=5, logit_scale=100.0)
full = eng.classify_same_size(crops)
print("same_size all151: max dlogit", (full.logits - ref_logits).abs().max().item())
for bs in (16, 64, 151):
    outs = []
    for i in range(0, 151, bs):
        r = eng.classify([c for c in crops[i:i + bs]])
        outs.append(r.logits.clone())
    lg = torch.cat(outs)
    d = (lg - ref_logits).abs().max(dim=1).values
    print(f"list path bs={bs}: max dlogit {d.max().item():.4f}; images over 0.02:", [(int(i), round(float(d[i]), 3)) for i in torch.nonzero(d > 0.02).flatten()])
    d2 = (lg - full.logits).abs().max(dim=1).values
    print(f"   vs same_size path: max {d2.max().item():.5f}", [(int(i), round(float(d2[i]), 3)) for i in torch.nonzero(d2 > 1e-3).flatten()])
