"""Parity of the CUDA hot path (through the C ABI) with outputs of the reference's own code (tests/golden/, made by
oracle/gen_golden.py from /root/reference/main.py on the repo's 150 dataset images + interior_sample.jpg).

Bars (BASELINE.json north_star / BASELINE.md section 4):
    embedding cosine similarity >= 0.999
    per-label logits (100 * cos, 437 labels) within 2e-2 absolute     <- 16-bit tensor-core path vs fp32 CPU reference
    identical top-1 style and top-5 label sets on >= 99 % of images; identical detector decision
These bars are asserted on the DEFAULT operand dtype - the one `Engine()`, `load()`, bench.py and smoke() run (fp16,
_lib.DEFAULT_OPERAND_DTYPE).  bf16 is an explicit non-default arm (`operand_dtype="bf16"`): `test_bf16_arm_*` below runs it
against what an 8-bit mantissa can deliver and records the numbers; it certifies nothing about the default path.

Top-5 SETS and near ties.  The fixture's weights are seeded, not pretrained, so the 299 "characteristics" logits of an
image are densely packed and rank 5 / rank 6 are often closer than the logit error itself (fp16: max 0.004).  Two
numbers are therefore reported: `same_top5_sets` (strict set equality of all five groups) and `same_top5_sets_tie_aware`,
where a swap only counts as a mismatch if the REFERENCE's own logits of the swapped labels differ from its rank-5 logit
by more than the stated logit tolerance (2e-2), i.e. the reference itself separates them.  The 99 % bar is asserted on
the tie-aware number; the strict number is asserted against STRICT_SETS and printed, with the worst swapped-label gap.
"""
import json

import numpy as np
import pytest
import torch

from _common import golden_json, golden_npz, label_layout, oracle_model, oracle_state_dict, top5_sets

pytestmark = pytest.mark.gpu

COS_BAR = 0.999
LOGIT_BAR = 2e-2
AGREE_BAR = 0.99
STRICT_SETS = 0.97   # strict set equality of all five groups, near-tie flips included (the 99 % bar is on the tie-aware count)


@pytest.fixture(scope="module")
def product(iic):
    """the product exactly as a user gets it: no operand dtype named -> the default"""
    model, preprocess = iic.load("ViT-B/16", device="cuda", state_dict=oracle_state_dict())
    model.mode = iic._lib.operand_dtype_name(model.visual.operand_dtype)
    assert model.mode == iic._lib.DEFAULT_OPERAND_DTYPE == "f16"
    return model, preprocess


@pytest.fixture(scope="module")
def product_bf16(iic):
    model, preprocess = iic.load("ViT-B/16", device="cuda", state_dict=oracle_state_dict(), operand_dtype="bf16")
    model.mode = "bf16"
    return model, preprocess


def _check(mode, cos, dl, style, sets, det):
    """north-star bars, no per-dtype exceptions"""
    sets, sets_tie = sets
    assert cos.min().item() >= COS_BAR
    assert style >= AGREE_BAR and det >= AGREE_BAR
    assert dl.max().item() <= LOGIT_BAR
    assert sets_tie >= AGREE_BAR and sets >= STRICT_SETS


def _report(name, emb, emb_ref, logits, logits_ref):
    cos = torch.nn.functional.cosine_similarity(emb.double(), emb_ref.double(), dim=-1)
    dl = (logits.double() - logits_ref.double()).abs()
    print(f"\n[{name}] cos min {cos.min():.6f} mean {cos.mean():.6f} | logit |d| max {dl.max():.4f} "
          f"p99.9 {dl.flatten().quantile(0.999):.4f} mean {dl.mean():.4f}")
    return cos, dl


def _dump(name, metrics):
    import os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    path = os.path.join(out, "parity_metrics.json")
    allm = json.load(open(path)) if os.path.exists(path) else {}
    allm[name] = metrics
    json.dump(allm, open(path, "w"), indent=1)


def _agreement(res, ref, lab):
    ref_top5 = json.loads(str(ref["top5"]))
    n = len(ref_top5)
    ti = res.topk_idx.cpu()
    same_style = same_sets = same_sets_tie = same_det = 0
    det_names = lab["detector"]
    ref_logits = ref["logits"]
    # global logit column of every (group, label)
    col, c0 = {}, len(det_names)
    for g in lab["group_order"]:
        for k, name in enumerate(lab["groups"][g]):
            col[(g, name)] = c0 + k
        c0 += len(lab["groups"][g])
    worst_gap = 0.0
    for i in range(n):
        got = top5_sets(ti[i], lab)
        want = {g: [l for l, _ in ref_top5[i][g]] for g in lab["group_order"]}
        same_style += got["styles"][0] == want["styles"][0]
        same_sets += all(set(got[g]) == set(want[g]) for g in lab["group_order"])
        tie_ok = True
        for g in lab["group_order"]:
            swapped = set(got[g]) ^ set(want[g])
            if swapped:
                thr = min(float(ref_logits[i, col[(g, l)]]) for l in want[g])       # the reference's rank-5 logit
                gap = max(abs(float(ref_logits[i, col[(g, l)]]) - thr) for l in swapped)
                worst_gap = max(worst_gap, gap)
                tie_ok = tie_ok and gap <= LOGIT_BAR
        same_sets_tie += tie_ok
        interior = float(res.split_sum[i, 0])
        non = float(res.probs[i, lab["n_interior"]:len(det_names)].sum())
        is_int = interior > non and float(res.topk_val[i, 0, 0]) > 0.3
        same_det += (is_int == bool(ref["det_is"][i])) and det_names[int(ti[i, 0, 0])] == str(ref["det_cat"][i])
    return same_style / n, (same_sets / n, same_sets_tie / n, worst_gap), same_det / n


def _run(product, iic, ref_name, vision_lora):
    model, _ = product
    lab, sizes, split = label_layout()
    text = torch.from_numpy(golden_npz("text_features.npz")["text"]).cuda()
    crops = torch.from_numpy(golden_npz("crops_u8.npz")["crops"]).cuda()
    ref = golden_npz(ref_name)
    eng = model.visual.sync_engine(use_lora=vision_lora)
    eng.set_labels(text, sizes, split, topk=5, logit_scale=100.0)
    eng._labels_owner = None
    res = eng.classify_same_size(crops)
    torch.cuda.synchronize()
    emb_ref = torch.from_numpy(ref["emb"]).cuda()
    logits_ref = torch.from_numpy(ref["logits"]).cuda()
    if vision_lora:
        # the reference's detector runs the un-LoRA'd tower: its 40 logits come from a second pass (main.py:238)
        eng0 = model.visual.sync_engine(use_lora=False)
        res0 = eng0.classify_same_size(crops)
        torch.cuda.synchronize()
        logits = torch.cat([res0.logits[:, :40], res.logits[:, 40:]], 1)
        res.probs[:, :40], res.split_sum[:, 0] = res0.probs[:, :40], res0.split_sum[:, 0]
        res.topk_val[:, 0], res.topk_idx[:, 0] = res0.topk_val[:, 0], res0.topk_idx[:, 0]
        model.visual.sync_engine(use_lora=True)
    else:
        logits = res.logits
    cos, dl = _report(ref_name, res.embedding, emb_ref, logits, logits_ref)
    style, (sets, sets_tie, worst_gap), det = _agreement(res, ref, lab)
    print(f"[{ref_name}] same top-1 style {style:.4f}  same top-5 sets {sets:.4f} (tie-aware {sets_tie:.4f}, worst reference "
          f"logit gap of a swapped label {worst_gap:.4f})  same detector result {det:.4f}")
    _dump(ref_name + ":" + model.mode, dict(cos_min=cos.min().item(), cos_mean=cos.mean().item(), logit_abs_max=dl.max().item(),
                         logit_abs_p999=dl.flatten().quantile(0.999).item(), logit_abs_rms=dl.pow(2).mean().sqrt().item(),
                         frac_logits_over_0p02=(dl > 2e-2).double().mean().item(), same_top1_style=style,
                         same_top5_sets=sets, same_top5_sets_tie_aware=sets_tie, worst_swapped_label_ref_gap=worst_gap,
                         same_detector=det, n_images=int(cos.numel())))
    return cos, dl, style, (sets, sets_tie), det


def test_dataset_parity_shipped_checkpoint(product, iic):
    """config 2: the 150 dataset images + interior_sample.jpg, shipped LoRA checkpoint (vision delta == 0, F7)."""
    cos, dl, style, sets, det = _run(product, iic, "ref_shipped.npz", vision_lora=False)
    _check(product[0].mode, cos, dl, style, sets, det)


def test_dataset_parity_vision_lora(product, iic):
    """same images with a seeded NON-zero LoRA on the vision MLPs: the fused LoRA k-block must carry the delta, and
    an out_proj LoRA must stay without effect (F4)."""
    from oracle import ref_semantics as RS
    model, _ = product
    wrapped = iic.replace_linears_with_lora(model, rank=4, alpha=8)
    assert len(wrapped) == 72
    # seed the LoRA on an ORACLE model wrapped by the restated reference code, then copy by parameter name
    import copy
    omodel = copy.deepcopy(oracle_model())
    assert len(RS.replace_linears_with_lora(omodel, rank=4, alpha=8)) == 72
    RS.seed_vision_lora(omodel, seed=1234)
    src = {n: p for n, p in omodel.named_parameters() if n.startswith("visual.") and "lora" in n}
    assert len(src) == 72
    for n, p in model.named_parameters():
        if n in src:
            p.data = src[n].detach().to(p.device)
    ref_a, ref_b = golden_npz("ref_shipped.npz")["emb"], golden_npz("ref_visionlora.npz")["emb"]
    assert np.linalg.norm(ref_b - ref_a) / np.linalg.norm(ref_a) > 0.05  # the delta is large: cannot pass by skipping it
    cos, dl, style, sets, det = _run(product, iic, "ref_visionlora.npz", vision_lora=True)
    _check(product[0].mode, cos, dl, style, sets, det)


def test_encode_image_entry_point(product, iic):
    """model.encode_image(float CHW batch) - the call the reference makes (main.py:204, 444, 503)."""
    model, _ = product
    model.visual.sync_engine(use_lora=False)
    crops = torch.from_numpy(golden_npz("crops_u8.npz")["crops"][:16]).cuda()
    mean = torch.tensor([0.48145466, 0.4578275, 0.40821073], device="cuda").view(1, 3, 1, 1)
    std = torch.tensor([0.26862954, 0.26130258, 0.27577711], device="cuda").view(1, 3, 1, 1)
    x = (crops.permute(0, 3, 1, 2).float() / 255 - mean) / std
    saved = model.visual.apply_out_proj_lora
    eng = model.visual.engine()
    emb = eng.encode_image(x)
    ref = torch.from_numpy(golden_npz("ref_shipped.npz")["emb_det"][:16]).cuda()
    cos = torch.nn.functional.cosine_similarity(emb.double(), ref.double(), dim=-1)
    assert cos.min().item() >= COS_BAR, cos.min()
    assert emb.dtype == torch.float32 and emb.shape == (16, 512)
    model.visual.apply_out_proj_lora = saved


def test_batch_invariance_and_determinism(product, iic):
    """Size-independent properties of the path: an image's result does not depend on the batch it travels in (rows never mix:
    bit-exact for any batch split), a second run gives the same bits, and at the BASELINE batch size (1024 / GPU) a
    permutation of the inputs permutes the outputs."""
    model, _ = product
    lab, sizes, split = label_layout()
    text = torch.from_numpy(golden_npz("text_features.npz")["text"]).cuda()
    crops = torch.from_numpy(golden_npz("crops_u8.npz")["crops"]).cuda()
    eng = model.visual.sync_engine(use_lora=False)
    eng.set_labels(text, sizes, split, topk=5, logit_scale=100.0)
    eng._labels_owner = None
    full = eng.classify_same_size(crops)
    logits, emb, ti = full.logits.clone(), full.embedding.clone(), full.topk_idx.clone()
    again = eng.classify_same_size(crops)
    assert torch.equal(again.logits, logits) and torch.equal(again.embedding, emb)
    for chunk in (1, 7, 64):
        parts = [eng.classify_same_size(crops[i:i + chunk]) for i in range(0, min(crops.shape[0], 3 * chunk), chunk)]
        got = torch.cat([p.logits for p in parts], 0)
        assert torch.equal(got, logits[:got.shape[0]]), chunk
    # full size: 1024 images (the 151 crops tiled), a random permutation
    g = torch.Generator(device="cuda").manual_seed(3)
    idx = torch.randint(0, crops.shape[0], (1024,), device="cuda", generator=g)
    big = crops[idx]
    r1 = eng.classify_same_size(big)
    l1 = r1.logits.clone()
    assert torch.equal(l1, logits[idx])
    perm = torch.randperm(1024, device="cuda", generator=g)
    r2 = eng.classify_same_size(big[perm])
    assert torch.equal(r2.logits, l1[perm]) and torch.equal(r2.topk_idx, ti[idx][perm])


def test_bf16_arm_dataset_parity(product_bf16, iic):
    """NON-DEFAULT arm (operand_dtype="bf16"), recorded for comparison: an 8-bit mantissa on the GEMM A operands alone costs
    ~0.026 on a 100 * cos logit (tools/error_budget.py emulates it inside the fp32 oracle: LayerNorm outputs 0.019, GELU
    outputs 0.016, attention outputs 0.011, QKV 0.006), so this arm is held to what bf16 can deliver - embedding cosine and
    top-1 / detector agreement at the north-star values, logits within 4e-2 - and NOT claimed to meet the 2e-2 logit bar."""
    cos, dl, style, (sets, sets_tie), det = _run(product_bf16, iic, "ref_shipped.npz", vision_lora=False)
    assert cos.min().item() >= COS_BAR and style >= AGREE_BAR and det >= AGREE_BAR
    assert dl.max().item() <= 4e-2 and sets_tie >= 0.95
