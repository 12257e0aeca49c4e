import os, sys, time, io, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
os.environ.setdefault("IIC_ALLOW_RANDOM_INIT", "1")
import iic_b200
from importlib import import_module
from bench_ingest import photo
from PIL import Image
jp = import_module("ai-interior-image-classifier_b200.jpeg")
clipc = import_module("ai-interior-image-classifier_b200.clip_compat")
dev = torch.device("cuda:0")
vis = clipc.build_visual("ViT-B/16", seed=0).cuda(); eng = vis.sync_engine()
groups = [40, 20, 12, 299, 36, 30]
eng.set_labels(torch.nn.functional.normalize(torch.randn(sum(groups), 512), dim=-1).cuda(), groups, [11, 0, 0, 0, 0, 0], topk=5, logit_scale=100.0)
rng = np.random.default_rng(0); uniq = []
for k in range(16):
    buf = io.BytesIO(); Image.fromarray(photo(rng, 768, 1024)).save(buf, "JPEG", quality=85, subsampling=2); uniq.append(buf.getvalue())
n = 1024; files = [uniq[i % 16] for i in range(n)]
side = torch.cuda.Stream(device=dev); main = torch.cuda.current_stream(dev)
def dec_side():
    with torch.cuda.stream(side):
        ims = jp.decode_jpeg_bytes(files, dev)[0]
        ev = torch.cuda.Event(); ev.record(side)
    return ims, ev
def run(overlap, nb=6):
    torch.cuda.synchronize(); t0 = time.perf_counter(); last = None
    nxt = dec_side() if overlap else None
    for b in range(nb):
        if overlap:
            ims, ev = nxt; main.wait_event(ev)
            nxt = dec_side() if b + 1 < nb else None
        else:
            ims = jp.decode_jpeg_bytes(files, dev)[0]
        r = eng.classify(ims, want_embedding=False)
        if last is not None: last.topk_idx.cpu()
        last = r
    last.topk_idx.cpu(); torch.cuda.synchronize()
    return nb * n / (time.perf_counter() - t0)
# pieces
def t(fn, k=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3
ims = jp.decode_jpeg_bytes(files, dev)[0]
print("decode ms", t(lambda: jp.decode_jpeg_bytes(files, dev)))
print("preprocess ms", t(lambda: eng.preprocess(ims)))
pt = eng.preprocess(ims)
print("encode+head ms", t(lambda: eng.classify_patches(pt, n, False)))
for ov in (False, True, False, True):
    run(ov, 2); print("overlap" if ov else "serial ", round(run(ov)), "img/s")
