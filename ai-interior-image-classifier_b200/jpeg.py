"""Host side of the GPU JPEG decoder (csrc/jpeg.cu, include/iic.h `iic_jpeg_*`): SURVEY 8(f) row N2, image ingest.

The reference loads every image with `Image.open(path).convert("RGB")` on a 4-thread pool (/root/reference/main.py:322-346,
404-417) and hands PIL objects to `preprocess`.  Here the FILE BYTES go to the device and are decoded there, bit-identically
to Pillow, into uint8 HWC tensors that `iic_preprocess` reads in place:

    decode_jpeg_bytes(list_of_bytes, device) -> ([uint8 [H, W, 3] cuda tensor | None, ...], [reason | "", ...])
    decode_jpeg_files(paths, device)         -> the same, reading the files straight into one pinned buffer

`None` marks a file outside the decoder's envelope (progressive, CMYK, ... - `reason` says which); the caller keeps those on
the host path, as it does for PNG files and URLs.  There is no CPU fallback in here: without the CUDA library this raises.
"""
from __future__ import annotations

import ctypes as C
import os
from collections import deque
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L

JPEG_OK, JPEG_UNSUPPORTED, JPEG_CORRUPT = 0, 1, 2
_PAD = 64          # the device blob must be readable 8 bytes past its end (include/iic.h); keeps the next buffer aligned too
_inflight: deque = deque()   # (event, tensors kept alive until the stream has consumed them)


def _retire(keep: int = 4) -> None:
    while len(_inflight) > keep:
        ev, _ = _inflight.popleft()
        ev.synchronize()


class JpegPlan:
    """Header parse + device layout of one batch of files (host only, no CUDA call)."""

    def __init__(self, blob: torch.Tensor, offsets: np.ndarray):
        self.lib = L.load()
        self.n = len(offsets) - 1
        self._offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        h = C.c_void_p()
        rc = self.lib.iic_jpeg_plan_create(blob.data_ptr(), self._offsets.ctypes.data_as(C.POINTER(C.c_int64)), self.n, C.byref(h))
        if rc != L.IIC_OK:
            raise RuntimeError(f"iic_jpeg_plan_create failed (code {rc})")
        self.h = h
        w, hh, st = C.c_int(), C.c_int(), C.c_int()
        self.sizes: List[Tuple[int, int]] = []       # (height, width), (0, 0) outside the envelope
        self.status: List[int] = []
        self.reasons: List[str] = []
        for i in range(self.n):
            self.lib.iic_jpeg_plan_info(h, i, C.byref(w), C.byref(hh), C.byref(st))
            self.sizes.append((hh.value, w.value))
            self.status.append(st.value)
            self.reasons.append("" if st.value == JPEG_OK else (self.lib.iic_jpeg_plan_reason(h, i) or b"").decode())
        self.staging_bytes = int(self.lib.iic_jpeg_plan_staging_bytes(h))
        self.scratch_bytes = int(self.lib.iic_jpeg_plan_scratch_bytes(h))

    def close(self) -> None:
        if self.h:
            self.lib.iic_jpeg_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def _decode(blob: torch.Tensor, offsets: np.ndarray, device) -> Tuple[List[Optional[torch.Tensor]], List[str]]:
    """blob: pinned uint8 host tensor holding the files back to back (+ _PAD spare bytes)."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("the JPEG decoder runs on a CUDA device only (no CPU path)")
    plan = JpegPlan(blob, offsets)
    try:
        ok = [i for i in range(plan.n) if plan.status[i] == JPEG_OK]
        out: List[Optional[torch.Tensor]] = [None] * plan.n
        if not ok:
            return out, plan.reasons
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device)
            dev_blob = blob.to(device, non_blocking=True)
            # one allocation for all decoded images of the batch (256-byte aligned slices)
            starts, total = [], 0
            for i in ok:
                starts.append(total)
                total += (plan.sizes[i][0] * plan.sizes[i][1] * 3 + 255) // 256 * 256
            pixels = torch.empty(total, dtype=torch.uint8, device=device)
            ptrs = (C.c_void_p * plan.n)()
            for i, s0 in zip(ok, starts):
                h, w = plan.sizes[i]
                out[i] = pixels[s0:s0 + h * w * 3].view(h, w, 3)
                ptrs[i] = pixels.data_ptr() + s0
            staging = torch.empty(max(plan.staging_bytes, 16), dtype=torch.uint8, pin_memory=True)
            scratch = torch.empty(plan.scratch_bytes, dtype=torch.uint8, device=device)
            rc = plan.lib.iic_jpeg_decode(plan.h, dev_blob.data_ptr(), ptrs, staging.data_ptr(), scratch.data_ptr(), stream.cuda_stream)
            if rc != L.IIC_OK:
                raise RuntimeError(f"iic_jpeg_decode failed (code {rc})")
            ev = torch.cuda.Event()
            ev.record(stream)
            _inflight.append((ev, (blob, staging, dev_blob, scratch)))   # stream-ordered consumers: keep alive until done
            _retire()
        return out, plan.reasons
    finally:
        plan.close()


def _pinned(nbytes: int) -> torch.Tensor:
    return torch.empty(nbytes + _PAD, dtype=torch.uint8, pin_memory=True)


def decode_jpeg_bytes(files: Sequence[bytes], device) -> Tuple[List[Optional[torch.Tensor]], List[str]]:
    offsets = np.zeros(len(files) + 1, dtype=np.int64)
    np.cumsum([len(b) for b in files], out=offsets[1:])
    blob = _pinned(int(offsets[-1]))
    view = blob.numpy()
    for b, lo, hi in zip(files, offsets[:-1], offsets[1:]):
        view[lo:hi] = np.frombuffer(b, dtype=np.uint8)
    view[offsets[-1]:] = 0
    return _decode(blob, offsets, device)


def decode_jpeg_files(paths: Sequence[str], device) -> Tuple[List[Optional[torch.Tensor]], List[str]]:
    """Reads the files straight into ONE pinned buffer (no per-file bytes objects) and decodes them on `device`.
    An unreadable file gets (None, "<error>") like a file outside the envelope."""
    sizes, errs = [], [""] * len(paths)
    for i, p in enumerate(paths):
        try:
            sizes.append(os.path.getsize(p))
        except OSError as e:
            sizes.append(0)
            errs[i] = f"unreadable: {e}"
    offsets = np.zeros(len(paths) + 1, dtype=np.int64)
    np.cumsum(sizes, out=offsets[1:])
    blob = _pinned(int(offsets[-1]))
    view = memoryview(blob.numpy())
    for i, p in enumerate(paths):
        if errs[i]:
            continue
        try:
            with open(p, "rb", buffering=0) as f:
                got = f.readinto(view[offsets[i]:offsets[i + 1]])
            if got != sizes[i]:
                errs[i] = "short read"
                view[offsets[i]:offsets[i] + 2] = b"\0\0"
        except OSError as e:
            errs[i] = f"unreadable: {e}"
    view[offsets[-1]:] = bytes(_PAD)
    imgs, reasons = _decode(blob, offsets, device)
    return imgs, [e or r for e, r in zip(errs, reasons)]
