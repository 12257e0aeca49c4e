"""Measurements for the SURVEY 8(f) "next" rows that sit either side of the image hot path:
  N3  text tower: label-matrix encode (437 prompts x 77 tokens, ViT-B/16 text tower with a live rank-4 LoRA) on the engine
      vs the PyTorch module on the same GPU and on the host CPU (what the reference does at start-up, main.py:296-311);
  N2  ingest: JPEG decode of a batch of files by nvJPEG on the device vs PIL on the host thread pool (main.py:345-346),
      measured up to "uint8 pixels resident on the GPU".
Prints one JSON line."""
import io, json, os, sys, tempfile, time
os.environ.setdefault("IIC_ALLOW_RANDOM_INIT", "1")       # synthetic benchmark: seeded weights by design
os.environ.setdefault("IIC_ALLOW_STANDIN_TOKENIZER", "1")
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import iic_b200
from importlib import import_module
an = import_module("ai-interior-image-classifier_b200.analyzer")
clipc = import_module("ai-interior-image-classifier_b200.clip_compat")


def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


out = {}
# ---- N3 ----
model, _ = iic_b200.load("ViT-B/16", device="cuda")
iic_b200.replace_linears_with_lora(model, rank=4, alpha=8)
for n, p in model.named_parameters():
    if n.startswith("transformer.") and n.endswith("lora_B"):
        p.data.normal_(0, 0.01)
tok = clipc.tokenize([f"wnętrze z etykietą numer {i}" for i in range(437)]).cuda()
with torch.no_grad():
    model.text_on_engine = False
    t_torch = timeit(lambda: model.encode_text(tok))
    ref = model.encode_text(tok)
    model.text_on_engine = True
    t_eng = timeit(lambda: model.encode_text(tok))
    got = model.encode_text(tok)
    cpu = iic_b200.load("ViT-B/16", device="cpu")[0]
    t0 = time.perf_counter(); cpu.encode_text(tok.cpu()[:64]); t_cpu = (time.perf_counter() - t0) * 437 / 64
cos = torch.nn.functional.cosine_similarity(got.double(), ref.double(), dim=-1).min().item()
out["text_tower_437_prompts"] = {"engine_ms": 1e3 * t_eng, "pytorch_fp32_same_gpu_ms": 1e3 * t_torch, "pytorch_fp32_host_cpu_ms_est": 1e3 * t_cpu,
                                 "host_cores": os.cpu_count(), "cos_min_vs_pytorch": cos}
# ---- N2 ----
from PIL import Image
crops = np.load(os.path.join(ROOT, "tests", "golden", "crops_u8.npz"))["crops"]
tmp = tempfile.mkdtemp()
paths = []
for i in range(256):
    im = Image.fromarray(crops[i % len(crops)]).resize((1024, 768))        # a typical dataset size
    p = os.path.join(tmp, f"{i}.jpg"); im.save(p, quality=90); paths.append(p)


def host():
    ims = an.load_images(paths, "cuda", gpu_decode=False)
    return [clipc.Preprocess._to_u8(im, torch.device("cuda")) for im in ims]


def dev():
    return [im.tensor for im in an.load_images(paths, "cuda", gpu_decode=True)]


t_host, t_dev = timeit(host, n=3, warm=1), timeit(dev, n=3, warm=1)
out["jpeg_ingest_256_files_1024x768"] = {"pil_host_4_threads_img_per_s": 256 / t_host, "nvjpeg_device_img_per_s": 256 / t_dev}
print(json.dumps(out))
