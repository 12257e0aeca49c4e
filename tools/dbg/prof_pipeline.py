import os, sys, time, io, cProfile, pstats, numpy as np, torch, tempfile
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
os.environ.setdefault("IIC_ALLOW_RANDOM_INIT", "1")
import iic_b200
from importlib import import_module
from bench_ingest import photo
from PIL import Image
jp = import_module("ai-interior-image-classifier_b200.jpeg")
an = import_module("ai-interior-image-classifier_b200.analyzer")
clipc = import_module("ai-interior-image-classifier_b200.clip_compat")
dev = torch.device("cuda:0")
vis = clipc.build_visual("ViT-B/16", seed=0).cuda(); eng = vis.sync_engine()
groups = [40, 20, 12, 299, 36, 30]
eng.set_labels(torch.nn.functional.normalize(torch.randn(sum(groups), 512), dim=-1).cuda(), groups, [11, 0, 0, 0, 0, 0], topk=5, logit_scale=100.0)
rng = np.random.default_rng(0); d = tempfile.mkdtemp(); paths = []
uniq = []
for k in range(16):
    buf = io.BytesIO(); Image.fromarray(photo(rng, 768, 1024)).save(buf, "JPEG", quality=85, subsampling=2); uniq.append(buf.getvalue())
for i in range(1024):
    p = os.path.join(d, f"{i}.jpg"); open(p, "wb").write(uniq[i % 16]); paths.append(p)
def one():
    ims = an.load_images(paths, dev, True)
    r = eng.classify([im.tensor for im in ims], want_embedding=False)
    return r.topk_idx.cpu()
one(); one()
torch.cuda.synchronize(); t0 = time.perf_counter(); one(); print("one chunk ms", (time.perf_counter() - t0) * 1e3)
pr = cProfile.Profile(); pr.enable(); one(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
