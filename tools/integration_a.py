"""INTEGRATION.md section A, proven on a B200: the reference's OWN classes running on `import iic_b200 as clip`.

    tools/run_integration_a.sh        (stages /root/reference/main.py into the git-ignored oracle/_ref/, runs this on the GPU box)

sys.modules["clip"] = iic_b200, then `import main` (the unmodified reference file) and its InteriorImageDetector /
CachedInteriorAnalyzer are constructed on "cuda" and asked to analyse dataset crops; results are compared with the
reference's own CPU outputs in tests/golden (same seeded weights).  Writes a log the repo keeps under profiles/."""
import json, os, sys, tempfile, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
os.environ["IIC_ALLOW_STANDIN_TOKENIZER"] = "1"     # no BPE vocabulary offline (the goldens were made with the same stand-in)
import numpy as np, torch
import iic_b200
from _common import GOLDEN, golden_npz, golden_json, oracle_state_dict
from PIL import Image

log = {"torch": torch.__version__, "device": torch.cuda.get_device_name(0)}
# the reference calls clip.load("ViT-B/16", device=...) with no weights argument: serve the oracle's seeded tensors as "the checkpoint"
ckpt_dir = tempfile.mkdtemp()
torch.save(oracle_state_dict(), os.path.join(ckpt_dir, "ViT-B-16.pt"))
_load = iic_b200.load
iic_b200.load = lambda name, device=None, **kw: _load(name, device=device, download_root=ckpt_dir, **kw)
sys.modules["clip"] = iic_b200
import main as ref_main                       # /root/reference/main.py, unmodified
log["reference_file"] = ref_main.__file__

work = tempfile.mkdtemp()
os.chdir(work)                                # main.py opens interior_dataset.json relative to the CWD (main.py:264)
with open("interior_dataset.json", "w", encoding="utf-8") as f:
    json.dump(golden_json("interior_dataset_fixture.json"), f, ensure_ascii=False)
crops = golden_npz("crops_u8.npz")
files = [str(f) for f in crops["files"]]
os.makedirs("dataset_images", exist_ok=True)
paths = []
for fn, c in list(zip(files, crops["crops"]))[:24]:
    p = os.path.splitext(fn)[0] + ".png"
    Image.fromarray(c).save(p)
    paths.append(p)
ref = golden_npz("ref_shipped.npz")
top5 = json.loads(str(ref["top5"]))

def compare(res, use_filter):
    same = 0
    for i, p in enumerate(paths):
        r = res[p]
        if use_filter and not bool(ref["det_is"][i]):
            ok = (not r["is_interior"]) and r["detected_category"] == str(ref["det_cat"][i])
        else:
            ok = r["is_interior"] and all([l for l, _ in r["analysis"][g]] == [l for l, _ in top5[i][g]] for g in top5[i])
        same += bool(ok)
    return same

for use_lora in (False, True):
    tag = f"use_lora={use_lora}"
    try:
        a = ref_main.CachedInteriorAnalyzer(use_lora=use_lora, lora_weights_path=None, lora_rank=4, lora_alpha=8, device="cuda")
        # the label matrix of the goldens was computed with the shipped text-LoRA checkpoint live; inject it (as the tests do)
        lab = golden_json("labels.json")
        text = torch.from_numpy(golden_npz("text_features.npz")["text"]).cuda()
        own_text_ok = all(torch.isfinite(v).all().item() for v in a.text_features_cache.values())
        a.detector.text_features = text[:40].clone()
        off = 40
        for g in lab["group_order"]:
            order = [lab["groups"][g].index(x) for x in a.all_categories[g]]      # the reference's own (set-iteration) label order
            a.text_features_cache[g] = text[off:off + len(order)][order].clone()
            off += len(order)
        r1 = a.analyze_images_batch(paths, batch_size=16, filter_interiors=True)
        r2 = a.analyze_images_batch(paths, batch_size=16, filter_interiors=False)
        det = a.detector.is_interior_image(Image.open(paths[0]))
        log[tag] = {"constructed": True, "own_label_features_finite": own_text_ok,
                    "engine_dtype": str(a.model.visual.engine().op_dtype),
                    "images_identical_to_reference_filter": f"{compare(r1, True)}/{len(paths)}",
                    "images_identical_to_reference_nofilter": f"{compare(r2, False)}/{len(paths)}",
                    "is_interior_image(first)": [bool(det[0]), float(det[1]), det[2]],
                    "reference_detector(first)": [bool(ref['det_is'][0]), float(ref['det_conf'][0]), str(ref['det_cat'][0])]}
    except Exception as e:  # noqa: BLE001
        log[tag] = {"constructed": False, "error": f"{type(e).__name__}: {e}", "trace": traceback.format_exc()[-1500:]}
out = os.path.join(ROOT, "gpurun_out", "r02_integration_a.json")
os.makedirs(os.path.dirname(out), exist_ok=True)
json.dump(log, open(out, "w"), indent=1, ensure_ascii=False)
print(json.dumps(log, indent=1, ensure_ascii=False))
