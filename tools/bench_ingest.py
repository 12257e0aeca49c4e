"""Row N2 (image ingest, /root/reference/main.py:322-346): JPEG files -> uint8 RGB pixels resident on the GPU.

  ours      jpeg.decode_jpeg_files: bytes read into one pinned buffer, H2D, Huffman + IDCT + upsample/colour kernels (csrc/jpeg.cu)
  pillow    the reference's way: Image.open(path).convert("RGB") on a thread pool (4 threads as main.py:345, and all host cores),
            then the H2D copy of the pixels
  nvjpeg    torchvision.io.decode_jpeg(device=cuda) on the same bytes (library call; round 1's opt-in path) when importable

Synthetic photos (structure + texture + noise, quality 85, 4:2:0) at 1024x768 unless --size; the per-kernel split of ours
comes from CUDA events around three extra decodes.  Writes one JSON object (stdout, and --out)."""
import argparse, io, json, os, sys, tempfile, time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import iic_b200  # noqa: E402
from importlib import import_module  # noqa: E402

jp = import_module("ai-interior-image-classifier_b200.jpeg")


def photo(rng, h, w):
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    base = np.stack([128 + 90 * np.sin(xx / 37.0 + yy / 91.0), 128 + 80 * np.cos(yy / 23.0), 255 * xx / max(w - 1, 1)], axis=2)
    tex = 40 * np.sin(xx[..., None] * np.array([0.9, 1.7, 2.9]) + yy[..., None] * np.array([2.1, 0.3, 1.1]))
    return np.clip(base + tex + rng.normal(0, 6, (h, w, 3)), 0, 255).astype(np.uint8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--uniq", type=int, default=32)
    ap.add_argument("--size", default="1024x768")
    ap.add_argument("--quality", type=int, default=85)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    from PIL import Image
    W, H = (int(v) for v in a.size.split("x"))
    rng = np.random.default_rng(0)
    d = tempfile.mkdtemp(prefix="iic_ingest_")
    uniq = []
    for k in range(a.uniq):
        buf = io.BytesIO()
        Image.fromarray(photo(rng, H, W)).save(buf, "JPEG", quality=a.quality, subsampling=2)
        uniq.append(buf.getvalue())
    paths = []
    for i in range(a.n):
        p = os.path.join(d, f"{i:05d}.jpg")
        open(p, "wb").write(uniq[i % a.uniq])
        paths.append(p)
    nbytes = sum(len(uniq[i % a.uniq]) for i in range(a.n))
    dev = torch.device("cuda:0")
    res = {"n_files": a.n, "size": a.size, "quality": a.quality, "subsampling": "4:2:0", "mean_file_kb": nbytes / a.n / 1024,
           "megapixels": a.n * W * H / 1e6}

    # ---- ours: files -> device pixels (host wall clock incl. file reads, header parse, H2D, kernels, final sync)
    def ours():
        imgs, _ = jp.decode_jpeg_files(paths, dev)
        torch.cuda.synchronize()
        return imgs
    imgs = ours()
    ref = np.asarray(Image.open(paths[0]).convert("RGB"))
    res["bit_exact_vs_pillow_sample"] = bool(np.array_equal(imgs[0].cpu().numpy(), ref))
    ts = []
    for _ in range(a.reps):
        t0 = time.perf_counter(); ours(); ts.append(time.perf_counter() - t0)
    res["ours"] = {"images_s": a.n / min(ts), "ms": min(ts) * 1e3, "all_ms": [round(t * 1e3, 1) for t in ts]}
    del imgs

    # ---- ours, device part only: bytes already in the slot's pinned buffer; CUDA events around the enqueued work
    files = [open(p, "rb").read() for p in paths]
    offsets = np.zeros(a.n + 1, dtype=np.int64); np.cumsum([len(f) for f in files], out=offsets[1:])
    nb = int(offsets[-1])
    joined = np.frombuffer(b"".join(files), dtype=np.uint8)
    hts, dts = [], []
    for _ in range(a.reps + 2):
        slot = jp._next_slot(dev)
        slot.host_blob(nb).numpy()[:nb] = joined
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(); jp._decode(slot, nb, offsets); e1.record()
        hts.append((time.perf_counter() - t0) * 1e3)
        torch.cuda.synchronize()
        dts.append(e0.elapsed_time(e1))
    t0 = time.perf_counter(); plan = jp.JpegPlan(slot.blob, offsets); res["plan_ms"] = (time.perf_counter() - t0) * 1e3; plan.close()
    res["ours_device"] = {"images_s": a.n / (min(dts) * 1e-3), "ms": min(dts), "host_enqueue_ms": min(hts),
                          "gbs_compressed": nbytes / (min(dts) * 1e-3) / 1e9, "gpix_s": a.n * W * H / (min(dts) * 1e-3) / 1e9}
    try:
        import re
        from torch.profiler import ProfilerActivity, profile
        slot = jp._next_slot(dev)
        slot.host_blob(nb).numpy()[:nb] = joined
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            jp._decode(slot, nb, offsets); torch.cuda.synchronize()
        res["kernels_ms"] = {(re.search(r"jpeg_\w+", e.key) or re.search(r"\w+", e.key)).group(0): round(e.device_time_total / 1e3, 3)
                             for e in prof.key_averages() if e.device_time_total > 0}
    except Exception as e:  # noqa: BLE001
        res["kernels_ms"] = {"error": str(e)}

    # ---- the reference's way
    def pil_one(p):
        return np.asarray(Image.open(p).convert("RGB"))
    for threads in (4, os.cpu_count() or 4):
        sub = paths[: min(a.n, 64 * threads)]
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(pil_one, sub[:threads]))
            t0 = time.perf_counter()
            arrs = list(ex.map(pil_one, sub))
            t_dec = time.perf_counter() - t0
        t0 = time.perf_counter()
        devs = [torch.from_numpy(x).to(dev, non_blocking=True) for x in arrs]
        torch.cuda.synchronize()
        t_h2d = time.perf_counter() - t0
        res[f"pillow_{threads}_threads"] = {"images_s": len(sub) / (t_dec + t_h2d), "decode_images_s": len(sub) / t_dec, "files": len(sub)}
        del devs, arrs
    try:
        from torchvision.io import ImageReadMode, decode_jpeg
        blobs = [torch.frombuffer(bytearray(f), dtype=torch.uint8) for f in files]
        decode_jpeg(blobs[:8], device=dev, mode=ImageReadMode.RGB); torch.cuda.synchronize()
        t0 = time.perf_counter(); out = decode_jpeg(blobs, device=dev, mode=ImageReadMode.RGB); torch.cuda.synchronize()
        res["nvjpeg_torchvision"] = {"images_s": a.n / (time.perf_counter() - t0)}
        del out
    except Exception as e:  # noqa: BLE001
        res["nvjpeg_torchvision"] = {"unavailable": str(e)[:200]}
    res["host_cores"] = os.cpu_count()
    s = json.dumps(res)
    print(s)
    if a.out:
        open(a.out, "w").write(s + "\n")


if __name__ == "__main__":
    main()
