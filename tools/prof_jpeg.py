"""Two decodes of 1024 synthetic 1024x768 JPEG files (for ncu: launch list and a full capture of the Huffman kernel)."""
import io, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import iic_b200  # noqa: F401
from importlib import import_module
from PIL import Image
from bench_ingest import photo
jp = import_module("ai-interior-image-classifier_b200.jpeg")
rng = np.random.default_rng(0)
uniq = []
for k in range(16):
    buf = io.BytesIO(); Image.fromarray(photo(rng, 768, 1024)).save(buf, "JPEG", quality=85, subsampling=2); uniq.append(buf.getvalue())
files = [uniq[i % 16] for i in range(int(os.environ.get("N", "1024")))]
for _ in range(2):
    imgs, _ = jp.decode_jpeg_bytes(files, "cuda:0")
    torch.cuda.synchronize()
print("ok", len(imgs), tuple(imgs[0].shape))
