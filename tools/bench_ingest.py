"""Row N2 (image ingest, /root/reference/main.py:322-346): JPEG files -> uint8 RGB pixels resident on the GPU, and on to top-k.

  ours        jpeg.decode_jpeg_files: bytes read into one pinned buffer, H2D, Huffman + IDCT + upsample/colour kernels (csrc/jpeg.cu)
  pillow      the reference's way: Image.open(path).convert("RGB") on a thread pool (4 threads as main.py:345, and all host cores),
              then the H2D copy of the pixels
  nvjpeg      torchvision.io.decode_jpeg(device=cuda) on the same bytes (library call; round 1's opt-in path) when importable
  files_to_topk   with an engine: decode -> iic_preprocess (Pillow-exact resize 1024x768 -> 224) -> encoder -> head -> top-k on the
              host, batch after batch with the decoder's two in-flight slots (the product pipeline from FILES)

Synthetic photos (structure + texture + noise, quality 85, 4:2:0) at 1024x768 unless --size; the per-kernel split of ours comes
from torch's profiler around one extra decode.  `measure()` is also bench.py's `ingest` leg.  Writes one JSON object."""
import argparse, io, json, os, re, shutil, sys, tempfile, time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import iic_b200  # noqa: E402,F401
from importlib import import_module  # noqa: E402

jp = import_module("ai-interior-image-classifier_b200.jpeg")


def photo(rng, h, w):
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    base = np.stack([128 + 90 * np.sin(xx / 37.0 + yy / 91.0), 128 + 80 * np.cos(yy / 23.0), 255 * xx / max(w - 1, 1)], axis=2)
    tex = 40 * np.sin(xx[..., None] * np.array([0.9, 1.7, 2.9]) + yy[..., None] * np.array([2.1, 0.3, 1.1]))
    return np.clip(base + tex + rng.normal(0, 6, (h, w, 3)), 0, 255).astype(np.uint8)


def measure(n=1024, uniq=32, size="1024x768", quality=85, reps=5, dev=None, with_pillow=True, with_nvjpeg=True, engine=None,
            pipeline_batches=4):
    from PIL import Image
    dev = dev or torch.device("cuda:0")
    W, H = (int(v) for v in size.split("x"))
    rng = np.random.default_rng(0)
    d = tempfile.mkdtemp(prefix="iic_ingest_")
    try:
        blobs = []
        for k in range(uniq):
            buf = io.BytesIO()
            Image.fromarray(photo(rng, H, W)).save(buf, "JPEG", quality=quality, subsampling=2)
            blobs.append(buf.getvalue())
        paths = []
        for i in range(n):
            p = os.path.join(d, f"{i:05d}.jpg")
            with open(p, "wb") as f:
                f.write(blobs[i % uniq])
            paths.append(p)
        nbytes = sum(len(blobs[i % uniq]) for i in range(n))
        res = {"n_files": n, "size": size, "quality": quality, "subsampling": "4:2:0", "mean_file_kb": nbytes / n / 1024,
               "megapixels": n * W * H / 1e6, "reference": "Image.open(path).convert('RGB'), main.py:322-346"}

        # ---- ours: files -> device pixels (host wall clock incl. file reads, header parse, H2D, kernels, final sync)
        def ours():
            imgs, _ = jp.decode_jpeg_files(paths, dev)
            torch.cuda.synchronize()
            return imgs
        imgs = ours()
        ref = np.asarray(Image.open(paths[0]).convert("RGB"))
        res["bit_exact_vs_pillow_sample"] = bool(np.array_equal(imgs[0].cpu().numpy(), ref))
        del imgs
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); ours(); ts.append(time.perf_counter() - t0)
        res["ours"] = {"images_s": n / min(ts), "ms": min(ts) * 1e3, "what": "one blocking call: file reads + header parse + H2D + kernels + sync"}
        # back-to-back calls, no sync in between: the two slots overlap the host side of call i+1 with the device side of call i
        t0 = time.perf_counter()
        for _ in range(pipeline_batches):
            jp.decode_jpeg_files(paths, dev)
        torch.cuda.synchronize()
        res["ours_pipelined"] = {"images_s": pipeline_batches * n / (time.perf_counter() - t0), "calls": pipeline_batches}

        # ---- ours, device part only: bytes already in the slot's pinned buffer; CUDA events around the enqueued work
        files = [blobs[i % uniq] for i in range(n)]
        offsets = np.zeros(n + 1, dtype=np.int64); np.cumsum([len(f) for f in files], out=offsets[1:])
        nb = int(offsets[-1])
        joined = np.frombuffer(b"".join(files), dtype=np.uint8)
        hts, dts = [], []
        for _ in range(reps + 2):
            slot = jp._next_slot(dev)
            slot.host_blob(nb).numpy()[:nb] = joined
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(); jp._decode(slot, nb, offsets); e1.record()
            slot.lock.release()
            hts.append((time.perf_counter() - t0) * 1e3)
            torch.cuda.synchronize()
            dts.append(e0.elapsed_time(e1))
        res["ours_device"] = {"images_s": n / (min(dts) * 1e-3), "ms": min(dts), "host_enqueue_ms": min(hts),
                              "gbs_compressed": nbytes / (min(dts) * 1e-3) / 1e9, "gpix_s": n * W * H / (min(dts) * 1e-3) / 1e9,
                              "what": "CUDA events around H2D of the file bytes + descriptor copy + the three kernels"}
        try:
            from torch.profiler import ProfilerActivity, profile
            slot = jp._next_slot(dev)
            slot.host_blob(nb).numpy()[:nb] = joined
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                jp._decode(slot, nb, offsets); torch.cuda.synchronize()
            slot.lock.release()
            res["kernels_ms"] = {(re.search(r"jpeg_\w+", e.key) or re.search(r"\w+", e.key)).group(0): round(e.device_time_total / 1e3, 3)
                                 for e in prof.key_averages() if e.device_time_total > 0}
            # algorithmic bytes of the two bandwidth-bound kernels: coefficients in + samples out; samples in + RGB out
            blocks = n * ((H + 15) // 16) * ((W + 15) // 16) * 6
            k = res["kernels_ms"]
            if k.get("jpeg_idct_kernel"):
                res["idct_gbs"] = blocks * (128 + 64) / (k["jpeg_idct_kernel"] * 1e-3) / 1e9
            if k.get("jpeg_color_kernel"):
                res["color_gbs"] = (blocks * 64 + n * W * H * 3) / (k["jpeg_color_kernel"] * 1e-3) / 1e9
        except Exception as e:  # noqa: BLE001
            res["kernels_ms"] = {"error": str(e)}

        # ---- files -> top-k through the engine (general-size preprocess + encoder + head), pipelined over batches
        if engine is not None:
            def to_topk():
                t0 = time.perf_counter()
                last = None
                for _ in range(pipeline_batches):
                    ims = jp.decode_jpeg_files(paths, dev)[0]
                    r = engine.classify(ims, want_embedding=False)
                    if last is not None:
                        last.topk_idx.cpu()
                    last = r
                last.topk_idx.cpu()
                return pipeline_batches * n / (time.perf_counter() - t0)
            to_topk()
            # (decoding batch i+1 on a side stream during the encode of batch i was measured and is SLOWER, 6.3k vs 7.2k img/s: the
            # serial Huffman chains lose issue slots to the GEMM warps they share the SMs with)
            res["files_to_topk"] = {"images_s": to_topk(), "batches": pipeline_batches,
                                    "what": "decode_jpeg_files -> Engine.classify (Pillow-exact resize to 224, encoder, head) -> top-k on the host"}
            # the analyzer's way for long lists (analyzer.iter_loaded): chunk i + 1 is read, parsed and enqueued by a worker thread
            # while chunk i is preprocessed, encoded and scored
            an = import_module("ai-interior-image-classifier_b200.analyzer")
            many = paths * pipeline_batches

            def chunked():
                t0 = time.perf_counter()
                last = None
                for part, ims in an.iter_loaded(many, dev, True, chunk=n):
                    r = engine.classify([im.tensor for im in ims], want_embedding=False)
                    if last is not None:
                        last.topk_idx.cpu()
                    last = r
                last.topk_idx.cpu()
                return len(many) / (time.perf_counter() - t0)
            chunked()
            res["files_to_topk_overlapped"] = {"images_s": chunked(), "files": len(many), "chunk": n,
                                               "what": "analyzer.iter_loaded (the next chunk is read, parsed and enqueued by a worker thread) -> Engine.classify -> top-k on the host"}
            sub = paths[: min(n, 256)]

            def host_decoder():
                with ThreadPoolExecutor(max_workers=4) as ex:
                    arrs = list(ex.map(lambda p: np.array(Image.open(p).convert("RGB")), sub))
                return [torch.from_numpy(x).to(dev, non_blocking=True) for x in arrs]
            t0 = time.perf_counter()
            engine.classify(host_decoder(), want_embedding=False).topk_idx.cpu()
            res["files_to_topk_pillow_4_threads"] = {"images_s": len(sub) / (time.perf_counter() - t0), "files": len(sub)}

        # ---- the reference's way
        if with_pillow:
            def pil_one(p):
                return np.array(Image.open(p).convert("RGB"))
            for threads in sorted({4, os.cpu_count() or 4}):
                sub = paths[: min(n, 64 * threads)]
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
        if with_nvjpeg:
            try:
                from torchvision.io import ImageReadMode, decode_jpeg
                tb = [torch.frombuffer(bytearray(f), dtype=torch.uint8) for f in files[: min(n, 256)]]
                decode_jpeg(tb[:8], device=dev, mode=ImageReadMode.RGB); torch.cuda.synchronize()
                t0 = time.perf_counter(); out = decode_jpeg(tb, device=dev, mode=ImageReadMode.RGB); torch.cuda.synchronize()
                res["nvjpeg_torchvision"] = {"images_s": len(tb) / (time.perf_counter() - t0), "files": len(tb)}
                del out
            except Exception as e:  # noqa: BLE001
                res["nvjpeg_torchvision"] = {"unavailable": str(e)[:200]}
        res["host_cores"] = os.cpu_count()
        return res
    finally:
        shutil.rmtree(d, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--uniq", type=int, default=32)
    ap.add_argument("--size", default="1024x768")
    ap.add_argument("--quality", type=int, default=85)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--engine", action="store_true", help="also time files -> top-k through a seeded ViT-B/16 engine")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    eng = None
    if a.engine:
        os.environ.setdefault("IIC_ALLOW_RANDOM_INIT", "1")
        clipc = import_module("ai-interior-image-classifier_b200.clip_compat")
        vis = clipc.build_visual("ViT-B/16", seed=0).cuda()
        eng = vis.sync_engine()
        groups = [40, 20, 12, 299, 36, 30]
        eng.set_labels(torch.nn.functional.normalize(torch.randn(sum(groups), 512), dim=-1).cuda(), groups, [11, 0, 0, 0, 0, 0], topk=5,
                       logit_scale=100.0)
    res = measure(a.n, a.uniq, a.size, a.quality, a.reps, engine=eng)
    s = json.dumps(res)
    print(s)
    if a.out:
        open(a.out, "w").write(s + "\n")


if __name__ == "__main__":
    main()
