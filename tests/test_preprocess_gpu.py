"""Preprocess kernel vs the REAL reference pipeline (`clip._transform` = torchvision + Pillow, importable on the box):
bit-exact in fp32 CHW for every image size that occurs in the repo's dataset (synthetic pixels) and for real photos
(tests/golden/raw_subset.npz), and exact-to-rounding in the 16-bit patch-matrix layout the encoder consumes.
Integer/byte work (the resample) must be bit-exact; the float normalisation is IEEE fp32 and is also bit-exact.
"""
import numpy as np
import pytest
import torch

from _common import golden_json, golden_npz

pytestmark = pytest.mark.gpu


def _reference(arr):
    from PIL import Image
    from oracle import clip_ref
    return clip_ref.transform(224)(Image.fromarray(arr))


def _run(engine, arrays, L):
    dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrays]
    return engine.preprocess(dev, layout=L.OUT_CHW_F32)


def test_resize_crop_normalise_every_dataset_size(engine, iic):
    L = iic._lib
    sizes = golden_json("sizes.json")["sizes_wh"]
    rng = np.random.default_rng(0)
    # all 55 distinct (W, H) of the dataset + edge cases: already 224 on one side, exact 224x224, tall, tiny, huge
    extra = [(224, 224), (224, 300), (500, 224), (225, 224), (224, 225), (230, 1000), (3000, 2600), (224, 4000), (449, 449)]
    todo = [tuple(s) for s in sizes] + extra
    arrays = [rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8) for (w, h) in todo]
    out = _run(engine, arrays, L).cpu()
    bad = []
    for i, a in enumerate(arrays):
        ref = _reference(a)
        if not torch.equal(out[i], ref):
            bad.append((todo[i], (out[i] - ref).abs().max().item(), int((out[i] != ref).sum())))
    assert not bad, f"{len(bad)} sizes differ from Pillow: {bad[:5]}"


def test_resize_real_photos(engine, iic):
    L = iic._lib
    raw = golden_npz("raw_subset.npz")
    arrays = [raw[k] for k in sorted(raw) if k.startswith("img")]
    assert len(arrays) >= 5
    out = _run(engine, arrays, L).cpu()
    for i, a in enumerate(arrays):
        assert torch.equal(out[i], _reference(a)), (i, a.shape)
    # smooth gradients and constant images exercise clamping / rounding differently from noise
    h, w = 479, 640
    yy, xx = np.mgrid[0:h, 0:w]
    grad = np.stack([(xx * 255 // (w - 1)), (yy * 255 // (h - 1)), ((xx + yy) % 256)], -1).astype(np.uint8)
    flat = np.full((300, 260, 3), 255, np.uint8)
    zero = np.zeros((260, 300, 3), np.uint8)
    out = _run(engine, [grad, flat, zero], L).cpu()
    for o, a in zip(out, [grad, flat, zero]):
        assert torch.equal(o, _reference(a))


@pytest.mark.parametrize("mode", ["bf16", "f16"])
def test_patch_matrix_layout(iic, engine, engine_f16, mode):
    """patch-matrix output == rounded CHW output re-indexed as conv1's im2col (column = c*256 + ky*16 + kx)."""
    eng = engine if mode == "bf16" else engine_f16
    L = iic._lib
    crops = torch.from_numpy(golden_npz("crops_u8.npz")["crops"][:5]).cuda()
    chw = eng.preprocess_same_size(crops, layout=L.OUT_CHW_F32)
    ref = torch.stack([_reference(c) for c in crops.cpu().numpy()])
    assert torch.equal(chw.cpu(), ref)           # same-size fast path == Pillow (resize is the identity at 224x224)
    patches = eng.preprocess_same_size(crops).clone()
    want = chw.view(5, 3, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(5 * 196, 768).to(eng.op_dtype)
    assert torch.equal(patches, want)
    # general path, patch layout, and the float-CHW -> patch kernel behind encode_image
    arrays = [c.contiguous() for c in crops]
    assert torch.equal(eng.preprocess(arrays).clone(), want)
    assert torch.equal(eng.patchify(chw).clone(), want)


def test_empty_and_errors(engine, iic):
    assert engine.preprocess([], layout=iic._lib.OUT_CHW_F32).shape[0] == 0
