"""GPU JPEG decoder (csrc/jpeg.cu through the C ABI `iic_jpeg_*`) against Pillow - the reference's real decoder
(`Image.open(path).convert("RGB")`, /root/reference/main.py:330-334), present on the GPU box - and against the committed
golden vectors.  Bit-exact: integer / byte work."""
import io
import json
import os

import numpy as np
import pytest
import torch

from _common import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def jp(iic):
    from importlib import import_module
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return import_module("ai-interior-image-classifier_b200.jpeg")


def _pil(data):
    from PIL import Image
    return np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))


def _encode(arr, **kw):
    from PIL import Image, ImageFile
    ImageFile.MAXBLOCK = max(ImageFile.MAXBLOCK, 4 * arr.size)     # optimize=True writes the whole scan in one piece
    buf = io.BytesIO()
    Image.fromarray(arr).save(buf, "JPEG", **kw)
    return buf.getvalue()


def _photo(rng, h, w):
    """smooth structure + texture + noise: realistic coefficient statistics (long and short Huffman codes, EOB runs, ZRL)"""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    base = np.stack([128 + 90 * np.sin(xx / 37.0 + yy / 91.0), 128 + 80 * np.cos(yy / 23.0), 255 * xx / max(w - 1, 1)], axis=2)
    tex = 40 * np.sin(xx[..., None] * np.array([0.9, 1.7, 2.9]) + yy[..., None] * np.array([2.1, 0.3, 1.1]))
    return np.clip(base + tex + rng.normal(0, 12, (h, w, 3)), 0, 255).astype(np.uint8)


def test_fixtures_bit_exact(jp):
    meta = json.load(open(os.path.join(GOLDEN, "jpeg", "meta.json")))["files"]
    exp = np.load(os.path.join(GOLDEN, "jpeg", "expected.npz"))
    names = sorted(meta)
    imgs, reasons = jp.decode_jpeg_files([os.path.join(GOLDEN, "jpeg", n + ".jpg") for n in names], "cuda:0")
    for n, im, why in zip(names, imgs, reasons):
        if meta[n]["in_envelope"]:
            assert im is not None and why == "", (n, why)
            assert tuple(im.shape) == exp[n].shape and np.array_equal(im.cpu().numpy(), exp[n]), n
        else:
            assert im is None and why, n


@pytest.mark.parametrize("sub", [0, 1, 2], ids=["444", "422", "420"])
def test_fresh_encodes_bit_exact(jp, sub):
    """one batch of files of many sizes (1x1 up to 1024x768 and a 2000x1500 one), qualities, table kinds and restart intervals"""
    rng = np.random.default_rng(100 + sub)
    files = []
    for (h, w) in ((1, 1), (2, 5), (8, 8), (9, 17), (16, 3), (31, 33), (40, 4), (224, 224), (256, 256), (479, 640), (600, 800),
                   (768, 1024), (1500, 2000)):
        img = _photo(rng, h, w) if h * w > 4096 else rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        for kw in (dict(quality=85), dict(quality=20, optimize=True), dict(quality=98, restart_marker_blocks=5), dict(quality=100)):
            if h * w > 10 ** 6 and kw.get("quality") != 85:
                continue
            files.append(_encode(img, subsampling=sub, **kw))
    imgs, reasons = jp.decode_jpeg_bytes(files, "cuda:0")
    for i, (data, im) in enumerate(zip(files, imgs)):
        ref = _pil(data)
        assert im is not None, (i, reasons[i])
        got = im.cpu().numpy()
        assert got.shape == ref.shape and np.array_equal(got, ref), (i, ref.shape, int((got != ref).sum()))


def test_grayscale_and_mixed_batch(jp):
    from PIL import Image
    rng = np.random.default_rng(5)
    gray = io.BytesIO()
    Image.fromarray(_photo(rng, 300, 411)).convert("L").save(gray, "JPEG", quality=77)
    prog = _encode(_photo(rng, 64, 64), quality=85, progressive=True)
    ok = _encode(_photo(rng, 123, 77), quality=60, subsampling=2)
    files = [ok, prog, gray.getvalue(), b"garbage", ok]
    imgs, reasons = jp.decode_jpeg_bytes(files, "cuda:0")
    assert imgs[1] is None and "progressive" in reasons[1]
    assert imgs[3] is None and reasons[3]
    for i in (0, 2, 4):
        assert np.array_equal(imgs[i].cpu().numpy(), _pil(files[i])), i


def test_large_batch_equals_single_decodes(jp):
    """size-independent property at the bench's batch size: 1024 files decoded in one call == each file decoded alone"""
    rng = np.random.default_rng(9)
    uniq = [_encode(_photo(rng, 96 + 8 * k, 128 + 5 * k), quality=50 + 3 * k, subsampling=(k % 3)) for k in range(16)]
    singles = [jp.decode_jpeg_bytes([u], "cuda:0")[0][0].cpu().numpy() for u in uniq]
    for u, s in zip(uniq, singles):
        assert np.array_equal(s, _pil(u))
    files = [uniq[(7 * i) % 16] for i in range(1024)]
    imgs, _ = jp.decode_jpeg_bytes(files, "cuda:0")
    torch.cuda.synchronize()
    for i in range(0, 1024, 37):
        assert np.array_equal(imgs[i].cpu().numpy(), singles[(7 * i) % 16]), i
    stack = torch.stack([imgs[i] for i in range(0, 1024, 16)])          # all copies of uniq[0]
    assert bool((stack == stack[0]).all())


def test_truncated_and_corrupt_streams_do_not_hang(jp):
    """damaged entropy data: the decoder must come back (zeros are fed past the end, as libjpeg does) - no claim on the pixels"""
    rng = np.random.default_rng(11)
    good = _encode(_photo(rng, 200, 300), quality=80, subsampling=2)
    bad = bytearray(good)
    for k in range(700, len(bad) - 2, 97):
        bad[k] = (bad[k] * 31 + 7) & 0xFF if bad[k] != 0xFF else 0xFE
    files = [good[: len(good) // 2], bytes(bad), good[:-2], good]
    imgs, _ = jp.decode_jpeg_bytes(files, "cuda:0")
    torch.cuda.synchronize()
    assert all(im is not None and tuple(im.shape) == (200, 300, 3) for im in imgs)
    assert np.array_equal(imgs[3].cpu().numpy(), _pil(good))


def test_load_images_gpu_decode_feeds_the_analyzer_path(jp, tmp_path, monkeypatch):
    """analyzer.load_images(gpu_decode=True): baseline files come back as device images with Pillow's pixels, a progressive file
    and a PNG go through the host path (PIL objects), order preserved"""
    from importlib import import_module
    from PIL import Image
    an = import_module("ai-interior-image-classifier_b200.analyzer")
    rng = np.random.default_rng(13)
    paths = []
    for k, kw in enumerate((dict(quality=85, subsampling=2), dict(quality=85, progressive=True), dict(quality=92, subsampling=0))):
        p = str(tmp_path / f"f{k}.jpg")
        open(p, "wb").write(_encode(_photo(rng, 240 + 10 * k, 320), **kw))
        paths.append(p)
    png = str(tmp_path / "g.png")
    Image.fromarray(_photo(rng, 50, 60)).save(png)
    paths.append(png)
    monkeypatch.setattr(an, "GPU_DECODE_MIN_FILES", 1)     # (calls with a handful of files stay on the host pool by default)
    out = an.load_images(paths, device="cuda:0", gpu_decode=True)
    assert isinstance(out[0], an.DeviceImage) and isinstance(out[2], an.DeviceImage)
    assert not isinstance(out[1], an.DeviceImage) and not isinstance(out[3], an.DeviceImage)
    for i in (0, 2):
        assert np.array_equal(out[i].tensor.cpu().numpy(), np.asarray(Image.open(paths[i]).convert("RGB")))
    for i in (1, 3):
        assert np.array_equal(np.asarray(out[i]), np.asarray(Image.open(paths[i]).convert("RGB")))


@pytest.mark.parametrize("chains,window", [("1", "32"), ("2", "32"), ("4", "32"), ("8", "32"), ("4", "0"), ("8", "1"), ("4", "300"), ("8", "100")])
def test_chains_inside_an_image(jp, monkeypatch, chains, window):
    """Small batches split every image into several decoding chains that start in the middle of the entropy-coded segment in a
    guessed state and hand over where their stream positions meet (jpeg_huffman_kernel<true>).  Whatever the chain count, and
    whether the speculative chains synchronise in their window (32 MCUs), cannot (window 0 / 1: they publish an untrusted
    boundary, are told to stop and the predecessor decodes on) or never get to publish (window longer than their segment):
    Pillow's pixels, bit for bit."""
    monkeypatch.setenv("IIC_JPEG_CHAINS", chains)
    monkeypatch.setenv("IIC_JPEG_CHAIN_WINDOW", window)
    rng = np.random.default_rng(21)
    files = []
    for (h, w, sub, q) in ((768, 1024, 2, 85), (600, 800, 0, 92), (1500, 2000, 2, 75), (479, 640, 1, 60), (1024, 768, 2, 97),
                           (512, 512, 2, 30), (333, 517, 2, 88), (64, 64, 2, 85)):
        files.append(_encode(_photo(rng, h, w), quality=q, subsampling=sub, optimize=(q % 2 == 0)))
    files.append(_encode(_photo(rng, 700, 900), quality=85, subsampling=2, restart_marker_blocks=7))     # restart markers: one chain
    from PIL import Image
    gray = io.BytesIO()
    Image.fromarray(_photo(rng, 640, 960)).convert("L").save(gray, "JPEG", quality=80)
    files.append(gray.getvalue())
    imgs, reasons = jp.decode_jpeg_bytes(files, "cuda:0")
    torch.cuda.synchronize()
    for i, (data, im) in enumerate(zip(files, imgs)):
        ref = _pil(data)
        assert im is not None, (i, reasons[i])
        got = im.cpu().numpy()
        assert got.shape == ref.shape and np.array_equal(got, ref), (i, ref.shape, int((got != ref).sum()))


def test_chains_on_truncated_streams_do_not_hang(jp, monkeypatch):
    monkeypatch.setenv("IIC_JPEG_CHAINS", "8")
    rng = np.random.default_rng(23)
    good = _encode(_photo(rng, 768, 1024), quality=85, subsampling=2)
    bad = bytearray(good)
    for k in range(len(bad) // 3, len(bad) - 2, 997):
        bad[k] = (bad[k] * 31 + 7) & 0xFF if bad[k] != 0xFF else 0xFE
    files = [good[: len(good) // 2], bytes(bad), good[: len(good) * 7 // 8], good]
    imgs, _ = jp.decode_jpeg_bytes(files, "cuda:0")
    torch.cuda.synchronize()
    assert all(im is not None and tuple(im.shape) == (768, 1024, 3) for im in imgs)
    assert np.array_equal(imgs[3].cpu().numpy(), _pil(good))


@pytest.mark.parametrize("chains", ["1", "8"])
def test_crafted_streams_bit_exact(jp, monkeypatch, chains):
    """tests/jpeg_craft.py: valid files with what third-party encoders emit and Pillow's encoder does not - Huffman codes up to 16
    bits (the canonical-code walk behind the look-up tables), restart intervals of 1 / 3 / 7 MCUs with fill bytes before RSTn, 16-bit
    quantisation tables, tables redefined before the scan, luma on table id 1, component ids 0-2 / 'RGB' under JFIF, Adobe markers,
    ZRL runs, blocks filled to coefficient 63, large magnitudes; also big enough to be cut into 8 chains"""
    import jpeg_craft as C
    from test_oracle import CRAFTED
    monkeypatch.setenv("IIC_JPEG_CHAINS", chains)
    rng = np.random.default_rng(3)
    files = [C.random_case(rng, **kw) for kw in CRAFTED]
    files += [C.random_case(rng, width=512, height=512), C.random_case(rng, width=640, height=400, sampling=(1, 1)),
              C.random_case(rng, width=700, height=520, sampling=(2, 1), deep=False), C.random_case(rng, width=512, height=384, restart=5)]
    rgb = [C.random_case(rng, width=32, height=32, jfif=False, adobe=0), C.random_case(rng, width=32, height=32, jfif=False, comp_ids=(82, 71, 66))]
    imgs, reasons = jp.decode_jpeg_bytes(files + rgb, "cuda:0")
    torch.cuda.synchronize()
    for i, data in enumerate(files):
        assert imgs[i] is not None, (i, reasons[i])
        ref = _pil(data)
        got = imgs[i].cpu().numpy()
        assert got.shape == ref.shape and np.array_equal(got, ref), (i, ref.shape, int((got != ref).sum()))
    assert imgs[-1] is None and imgs[-2] is None and "RGB" in reasons[-1] and "RGB" in reasons[-2]


def test_random_crafted_batch(jp):
    """48 random valid streams (tests/jpeg_craft.py) in ONE batch: random sizes from 1x1 to 300x300, every sampling layout and
    grayscale, random restart intervals and fill bytes, deep or flat Huffman tables, 8 / 16-bit quantisation tables, random ids"""
    import jpeg_craft as C
    rng = np.random.default_rng(2024)
    files = []
    for i in range(48):
        w, h = int(rng.integers(1, 301)), int(rng.integers(1, 301))
        gray = bool(rng.integers(0, 6) == 0)
        kw = dict(width=w, height=h, gray=gray, deep=bool(rng.integers(0, 2)), dqt16=bool(rng.integers(0, 4) == 0))
        if not gray:
            kw["sampling"] = [(1, 1), (2, 1), (2, 2)][int(rng.integers(0, 3))]
            kw["comp_ids"] = [(1, 2, 3), (0, 1, 2), (10, 20, 30)][int(rng.integers(0, 3))]
            kw["table_ids"] = [((0, 0), (1, 1), (1, 1)), ((1, 1), (0, 0), (0, 0)), ((0, 1), (1, 0), (0, 0))][int(rng.integers(0, 3))]
        if rng.integers(0, 3) == 0:
            kw["restart"] = int(rng.integers(1, 9))
            kw["fill_before_rst"] = int(rng.integers(0, 3))
        files.append(C.random_case(rng, **kw))
    imgs, reasons = jp.decode_jpeg_bytes(files, "cuda:0")
    torch.cuda.synchronize()
    for i, data in enumerate(files):
        assert imgs[i] is not None, (i, reasons[i])
        ref = _pil(data)
        got = imgs[i].cpu().numpy()
        assert got.shape == ref.shape and np.array_equal(got, ref), (i, ref.shape, int((got != ref).sum()))
