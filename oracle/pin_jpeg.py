"""Pins oracle/jpeg_ref.py against Pillow (the reference's real decoder) on the reference's own dataset, in the build
container:  python -m oracle.pin_jpeg [max_files]   (/root/reference must be mounted).  Writes tests/golden/jpeg_pin.json:
per file the SHA-256 of the decoded RGB array, identical for both decoders, or the reason the file is outside the oracle's
envelope (progressive files stay on the host path in the product as well)."""
import glob, hashlib, io, json, os, sys, time

import numpy as np
from PIL import Image

from oracle import jpeg_ref as J


def main():
    files = sorted(glob.glob("/root/reference/**/*.jp*g", recursive=True), key=os.path.getsize)
    limit = int(sys.argv[1]) if len(sys.argv) > 1 else len(files)
    out, ok, skipped, bad = {}, 0, 0, 0
    t0 = time.time()
    for f in files[:limit]:
        data = open(f, "rb").read()
        ref = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
        name = os.path.relpath(f, "/root/reference")
        try:
            got = J.decode_rgb(data)
        except J.Unsupported as e:
            out[name] = {"unsupported": str(e)}
            skipped += 1
            continue
        same = got.shape == ref.shape and np.array_equal(got, ref)
        out[name] = {"shape": list(ref.shape), "sha256": hashlib.sha256(ref.tobytes()).hexdigest(), "oracle_equals_pillow": bool(same)}
        ok += same
        bad += not same
        print(f"{name} {ref.shape} {'OK' if same else 'MISMATCH'} ({time.time() - t0:.0f}s)", flush=True)
    summary = {"files": len(out), "bit_exact": ok, "mismatch": bad, "outside_envelope": skipped, "pillow": Image.__version__}
    json.dump({"summary": summary, "files": out}, open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "jpeg_pin.json"), "w"), indent=1)
    print(summary)


if __name__ == "__main__":
    main()
