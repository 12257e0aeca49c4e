"""Writes the JPEG fixtures under tests/golden/jpeg/: small files encoded by Pillow from crops of the reference's own photos
(tests/golden/raw_subset.npz) in every layout of the decoder's envelope, plus files outside it, and `expected.npz` = what
Pillow (`Image.open(...).convert("RGB")`, the reference's load_image, main.py:330-334) decodes them to.
    python -m oracle.gen_jpeg_fixtures"""
import io, json, os

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden", "jpeg")


def main():
    raw = np.load(os.path.join(HERE, "..", "tests", "golden", "raw_subset.npz"))
    a, b, c = raw["img0"], raw["img5"], raw["img2"]
    cases = {
        "q85_420_96x72": (a[100:172, 200:296], dict(quality=85, subsampling=2)),
        "q95_444_50x75": (b[300:375, 400:450], dict(quality=95, subsampling=0)),
        "q70_422_131x100": (c[50:150, 100:231], dict(quality=70, subsampling=1)),
        "q30_420_optimized_121x67": (b[100:167, 500:621], dict(quality=30, subsampling=2, optimize=True)),
        "q90_420_restart_80x64": (a[200:264, 300:380], dict(quality=90, subsampling=2, restart_marker_blocks=3)),
        "q100_444_17x23": (c[10:33, 20:37], dict(quality=100, subsampling=0)),
        "q75_420_1x33": (a[0:33, 5:6], dict(quality=75, subsampling=2)),
        "q80_gray_64x48": (np.asarray(Image.fromarray(b[200:248, 100:164]).convert("L")), dict(quality=80)),
        "q85_420_progressive_64x64": (a[0:64, 0:64], dict(quality=85, subsampling=2, progressive=True)),   # outside the envelope
        "q85_cmyk_32x32": (np.asarray(Image.fromarray(a[0:32, 0:32]).convert("CMYK")), dict(quality=85)),  # outside the envelope
    }
    expected, meta = {}, {}
    for name, (arr, kw) in cases.items():
        buf = io.BytesIO()
        (Image.fromarray(arr, "CMYK") if "cmyk" in name else Image.fromarray(arr)).save(buf, "JPEG", **kw)
        data = buf.getvalue()
        open(os.path.join(OUT, name + ".jpg"), "wb").write(data)
        im = Image.open(io.BytesIO(data))
        expected[name] = np.asarray(im.convert("RGB"))
        meta[name] = {"bytes": len(data), "mode": im.mode, "size": list(im.size), "progressive": bool(im.info.get("progressive")),
                      "in_envelope": not (im.info.get("progressive") or im.mode == "CMYK")}
    np.savez_compressed(os.path.join(OUT, "expected.npz"), **expected)
    json.dump({"pillow": Image.__version__, "files": meta}, open(os.path.join(OUT, "meta.json"), "w"), indent=1)
    print(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
