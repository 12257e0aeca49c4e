"""Throughput of the GENERAL preprocess path (Pillow-exact two-pass bicubic resize + centre crop + normalise, csrc/preprocess.cu
resize_h / resize_v kernels) on the dataset's real size mix (tests/golden/sizes.json: 151 images, 256x256 ... 2989x2592), images
resident on the device as uint8 HWC - BASELINE configs[1]'s preprocessing cost - next to the same-size fast path."""
import json, os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import iic_b200
eng = iic_b200.Engine(iic_b200.VIT_B_16, "cuda:0")
sizes = json.load(open(os.path.join(ROOT, "tests", "golden", "sizes.json")))["per_file"]
g = torch.Generator(device="cuda").manual_seed(0)
imgs = [torch.randint(0, 256, (h, w, 3), dtype=torch.uint8, device="cuda", generator=g) for (w, h) in sizes.values()]
in_bytes = sum(t.numel() for t in imgs)
out_bytes = len(imgs) * 196 * 768 * 2
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n, (time.perf_counter() - t0) / n * 1e3
dev_ms, wall_ms = timeit(lambda: eng.preprocess(imgs))
same = torch.randint(0, 256, (1024, 224, 224, 3), dtype=torch.uint8, device="cuda", generator=g)
fast_ms, _ = timeit(lambda: eng.preprocess_same_size(same), n=30)
print(json.dumps({
    "general_path": {"images": len(imgs), "megapixels_in": in_bytes / 3e6, "ms_device": dev_ms, "ms_wall_incl_host_tables": wall_ms,
                     "images_s": len(imgs) / (wall_ms * 1e-3), "input_gb_s_device": in_bytes / dev_ms / 1e6,
                     "note": "per call: host builds the per-image coefficient tables (cached per (in, out) size pair), two H2D copies of "
                             "descriptors/tables, resize_h + resize_v launches; wall clock includes that host work"},
    "same_size_fast_path": {"images": 1024, "ms_device": fast_ms, "images_s": 1024 / (fast_ms * 1e-3),
                            "gb_s": (same.numel() + 1024 * 196 * 768 * 2) / fast_ms / 1e6}}))
