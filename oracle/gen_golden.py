"""ORACLE tooling (run in the BUILD container only, where /root/reference is mounted):

    python oracle/gen_golden.py

Runs the reference's OWN code - /root/reference/main.py imported unmodified on top of the restated CLIP
(oracle/clip_ref.py, seeded weights, fp32, CPU) - over the repo's 150 dataset images + interior_sample.jpg and
writes the golden fixtures the GPU parity tests compare against (tests/golden/):

    crops_u8.npz          [151,224,224,3] uint8: Pillow Resize(224,BICUBIC)+CenterCrop(224) output of every image
                          (the bytes ToTensor/Normalize consume) + file names
    raw_subset.npz        decoded uint8 HWC originals of a few small images (resize-kernel parity on real photos)
    sizes.json            every distinct (W,H) in the dataset (resize-kernel parity on synthetic images of each)
    interior_dataset_fixture.json   the label schema in the reference's layout (paths relative)
    labels.json           label groups (sorted) + detector categories, row order of text_features
    text_features.npz     [437,512] fp32 label matrix: 40 detector prompts (base text tower) + 397 attribute prompts
                          (text tower WITH the shipped text-LoRA checkpoint live)
    ref_shipped.npz       reference outputs with lora_models/comprehensive_lora.pth: embeddings [151,512],
                          logits [151,437], detector triples, top-5 per group
    ref_visionlora.npz    same with a seeded NON-zero vision-tower LoRA (c_fc, c_proj live; out_proj set but dead, F4)
    meta.json             seed, versions, SHA-256 of every fixture, weight checksums, loader KATs (48 / 96)
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import clip_ref  # noqa: E402
from oracle import ref_semantics as RS  # noqa: E402


def sha256(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    os.makedirs(OUT, exist_ok=True)
    clip_ref.install_clip_stub()
    import main as ref_main  # the reference, unmodified
    from PIL import Image

    with open(os.path.join(REF, "interior_dataset.json"), encoding="utf-8") as f:
        dataset = json.load(f)
    files = []
    for item in dataset["training_data"]:
        if item["image_path"] not in files:
            files.append(item["image_path"])
    assert len(files) == 150 and len(dataset["training_data"]) == 151
    files.append("interior_sample.jpg")
    paths = [os.path.join(REF, p) for p in files]

    # ---------------------------------------------------------------- the reference analyzer (config 1 / 2)
    cwd = os.getcwd()
    os.chdir(REF)  # main.py opens interior_dataset.json relative to the CWD (main.py:264)
    try:
        t0 = time.time()
        analyzer = ref_main.CachedInteriorAnalyzer(use_lora=True, lora_weights_path="lora_models/comprehensive_lora.pth",
                                                   lora_rank=4, lora_alpha=8, device="cpu")
        print(f"reference analyzer built in {time.time() - t0:.1f}s")
    finally:
        os.chdir(cwd)
    kat = RS.loader_kat(ref_main, os.path.join(REF, "lora_models", "comprehensive_lora.pth"))
    print("loader KAT:", kat)

    # ---------------------------------------------------------------- preprocessing goldens
    imgs = [Image.open(p).convert("RGB") for p in paths]
    from torchvision.transforms import CenterCrop, InterpolationMode, Resize
    crop = [np.asarray(CenterCrop(224)(Resize(224, interpolation=InterpolationMode.BICUBIC)(im))) for im in imgs]
    crops = np.stack(crop).astype(np.uint8)
    np.savez_compressed(os.path.join(OUT, "crops_u8.npz"), crops=crops, files=np.array(files))
    sizes = sorted({im.size for im in imgs})
    with open(os.path.join(OUT, "sizes.json"), "w") as f:
        json.dump({"sizes_wh": sizes, "per_file": {fn: im.size for fn, im in zip(files, imgs)}}, f)
    small = [i for i, im in enumerate(imgs) if im.size[0] * im.size[1] <= 1024 * 768]
    picked, seen = [], set()
    for i in small:
        if imgs[i].size not in seen and len(picked) < 10:
            seen.add(imgs[i].size)
            picked.append(i)
    np.savez_compressed(os.path.join(OUT, "raw_subset.npz"), names=np.array([files[i] for i in picked]),
                        **{f"img{j}": np.asarray(imgs[i]) for j, i in enumerate(picked)})

    # ---------------------------------------------------------------- labels + text features
    groups = {g: sorted(analyzer.all_categories[g]) for g in RS.GROUP_ORDER}
    det_cats = list(analyzer.detector.categories)
    rows = [analyzer.detector.text_features]
    for g in RS.GROUP_ORDER:
        order = [analyzer.all_categories[g].index(lbl) for lbl in groups[g]]
        rows.append(analyzer.text_features_cache[g][order])
    text = torch.cat(rows, 0).contiguous()
    assert text.shape == (437, 512), text.shape
    np.savez_compressed(os.path.join(OUT, "text_features.npz"), text=text.numpy())
    with open(os.path.join(OUT, "labels.json"), "w", encoding="utf-8") as f:
        json.dump({"detector": det_cats, "n_interior": 11, "group_order": list(RS.GROUP_ORDER), "groups": groups}, f,
                  ensure_ascii=False)
    with open(os.path.join(OUT, "interior_dataset_fixture.json"), "w", encoding="utf-8") as f:
        json.dump(dataset, f, ensure_ascii=False, separators=(",", ":"))

    # ---------------------------------------------------------------- reference outputs
    def run_reference(tag: str):
        x = torch.stack([analyzer.preprocess(im) for im in imgs])
        # torchvision pipeline == crops (sanity of the crops fixture)
        chk = torch.from_numpy(crops[:4]).permute(0, 3, 1, 2).float().div(255)
        mean = torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(1, 3, 1, 1)
        std = torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(1, 3, 1, 1)
        assert torch.equal((chk - mean) / std, x[:4])
        emb, emb_det = [], []
        with torch.no_grad():
            for i in range(0, len(imgs), 16):
                emb.append(analyzer.model.encode_image(x[i:i + 16]))            # main.py:444
                emb_det.append(analyzer.detector.model.encode_image(x[i:i + 16]))  # main.py:204
        emb, emb_det = torch.cat(emb), torch.cat(emb_det)
        f = emb / emb.norm(dim=-1, keepdim=True)
        fd = emb_det / emb_det.norm(dim=-1, keepdim=True)
        logits = torch.cat([100.0 * fd @ text[:40].T, 100.0 * f @ text[40:].T], dim=1)
        # the reference's own entry points, image by image
        det, top5 = [], []
        for i, im in enumerate(imgs):
            det.append(analyzer.detector.is_interior_image(im, 0.3))                       # main.py:191-222
            res = analyzer._analyze_image_tensor_fast(x[i:i + 1])                          # main.py:500-510
            top5.append({g: [(lbl, float(p)) for lbl, p in res[g]] for g in RS.GROUP_ORDER})
        # the restated head (oracle/ref_semantics.py) must reproduce the reference's own outputs exactly
        for i in range(len(imgs)):
            d2 = RS.detector_decision(logits[i, :40], det_cats, 0.3)
            assert d2[0] == det[i][0] and abs(d2[1] - det[i][1]) < 1e-6 and d2[2] == det[i][2], (d2, det[i])
            t2 = RS.group_topk(logits[i, 40:], groups)
            for g in RS.GROUP_ORDER:
                assert [l for l, _ in t2[g]] == [l for l, _ in top5[i][g]], (i, g)
                assert np.allclose([p for _, p in t2[g]], [p for _, p in top5[i][g]], atol=1e-6)
        np.savez_compressed(os.path.join(OUT, f"ref_{tag}.npz"), emb=emb.numpy(), emb_det=emb_det.numpy(),
                            logits=logits.numpy(), det_is=np.array([d[0] for d in det]),
                            det_conf=np.array([d[1] for d in det], dtype=np.float64),
                            det_cat=np.array([d[2] for d in det]), top5=np.array(json.dumps(top5, ensure_ascii=False)))
        return emb, logits

    t0 = time.time()
    emb_a, logits_a = run_reference("shipped")
    print(f"reference pass (shipped checkpoint) {time.time() - t0:.1f}s")
    # batch entry point, both filter settings, must agree with the per-image path (main.py:371-469)
    os.chdir(REF)
    try:
        res_f = analyzer.analyze_images_batch(files[:12], batch_size=16, filter_interiors=True)
        res_n = analyzer.analyze_images_batch(files[:12], batch_size=16, filter_interiors=False)
    finally:
        os.chdir(cwd)
    with open(os.path.join(OUT, "ref_batch12.json"), "w", encoding="utf-8") as f:
        json.dump({"filter": res_f, "nofilter": res_n}, f, ensure_ascii=False)

    # vision LoRA made non-zero (seeded): the fused path cannot be skipped; out_proj LoRA set too but dead (F4)
    RS.seed_vision_lora(analyzer.model, seed=1234)
    t0 = time.time()
    emb_b, logits_b = run_reference("visionlora")
    print(f"reference pass (seeded vision LoRA) {time.time() - t0:.1f}s; "
          f"embedding change vs shipped: {((emb_b - emb_a).norm() / emb_a.norm()).item():.4f}")

    meta = {
        "generated_by": "oracle/gen_golden.py", "seed_model": 0, "seed_vision_lora": 1234,
        "torch": torch.__version__, "numpy": np.__version__, "pillow": __import__("PIL").__version__,
        "torchvision": __import__("torchvision").__version__,
        "weights_checksum": RS.weights_checksum(analyzer.detector.model),
        "loader_kat": kat, "n_images": len(files),
        "sha256": {fn: sha256(os.path.join(OUT, fn)) for fn in sorted(os.listdir(OUT)) if fn != "meta.json"},
    }
    with open(os.path.join(OUT, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(json.dumps({k: v for k, v in meta.items() if k != "sha256"}, indent=1))


if __name__ == "__main__":
    main()
