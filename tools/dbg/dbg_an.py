import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["IIC_ALLOW_RANDOM_INIT"] = "1"; os.environ["IIC_ALLOW_STANDIN_TOKENIZER"] = "1"
import numpy as np, torch, tempfile
from PIL import Image
import iic_b200
from _common import GOLDEN, golden_npz, golden_json, oracle_state_dict
crops = golden_npz("crops_u8.npz"); files = [str(f) for f in crops["files"]]
root = tempfile.mkdtemp(); os.makedirs(os.path.join(root, "dataset_images"))
for f, c in zip(files, crops["crops"]):
    Image.fromarray(c).save(os.path.join(root, os.path.splitext(f)[0] + ".png"))
model, pre = iic_b200.load("ViT-B/16", device="cuda", state_dict=oracle_state_dict())
a = iic_b200.CachedInteriorAnalyzer(use_lora=True, lora_weights_path=None, lora_rank=4, lora_alpha=8, device="cuda",
                                   json_path=os.path.join(GOLDEN, "interior_dataset_fixture.json"), model=model, preprocess=pre)
lab = golden_json("labels.json"); text = torch.from_numpy(golden_npz("text_features.npz")["text"]).cuda()
off = 40; a.detector.text_features = text[:40].clone()
for g in lab["group_order"]:
    n = len(lab["groups"][g]); a.text_features_cache[g] = text[off:off + n].clone(); off += n
ref = golden_npz("ref_shipped.npz"); top5 = json.loads(str(ref["top5"]))
paths = [os.path.join(root, os.path.splitext(f)[0] + ".png") for f in files]
for flt in (True, False):
    res = a.analyze_images_batch(paths, batch_size=16, filter_interiors=flt)
    for i, p in enumerate(paths):
        got = res[p]
        if not got["is_interior"]:
            continue
        for g, pairs in top5[i].items():
            gl = [l for l, _ in got["analysis"][g]]; wl = [l for l, _ in pairs]
            gp = [q for _, q in got["analysis"][g]]; wp = [q for _, q in pairs]
            if gl != wl or not np.allclose(gp, wp, atol=5e-3):
                print(flt, i, files[i], g, "\n   got ", list(zip(gl, np.round(gp, 4))), "\n   want", list(zip(wl, np.round(wp, 4))))
