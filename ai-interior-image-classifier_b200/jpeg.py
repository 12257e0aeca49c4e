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
import threading
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L

JPEG_OK, JPEG_UNSUPPORTED, JPEG_CORRUPT = 0, 1, 2
_PAD = 64          # the device blob must be readable 8 bytes past its end (include/iic.h); keeps the next buffer aligned too
_READ_THREADS = 8


class _Slot:
    """Grow-only buffers of one in-flight batch: pinned host (file bytes, descriptors) and device (file bytes, scratch).  Two
    slots per device alternate, so batch i+1 is read and parsed on the host while batch i is still being decoded; a slot is
    reused only after the event recorded behind its last decode has completed (cudaHostAlloc / cudaMalloc happen only when
    a buffer grows)."""

    def __init__(self, device: torch.device):
        self.device = device
        self.blob = torch.empty(0, dtype=torch.uint8)
        self.staging = torch.empty(0, dtype=torch.uint8)
        self.dev_blob = torch.empty(0, dtype=torch.uint8, device=device)
        self.scratch = torch.empty(0, dtype=torch.uint8, device=device)
        self.event: Optional[torch.cuda.Event] = None
        self.lock = threading.Lock()     # held from _next_slot() until the decode has been enqueued (several feeder threads)

    def wait(self) -> None:
        if self.event is not None:
            self.event.synchronize()
            self.event = None

    @staticmethod
    def _grow(t: torch.Tensor, nbytes: int, **kw) -> torch.Tensor:
        if t.numel() >= nbytes:
            return t
        return torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, **kw)

    def host_blob(self, nbytes: int) -> torch.Tensor:
        self.blob = self._grow(self.blob, nbytes + _PAD, pin_memory=True)
        return self.blob

    def reserve(self, blob_bytes: int, staging_bytes: int, scratch_bytes: int) -> None:
        self.staging = self._grow(self.staging, max(staging_bytes, 16), pin_memory=True)
        self.dev_blob = self._grow(self.dev_blob, blob_bytes + _PAD, device=self.device)
        self.scratch = self._grow(self.scratch, scratch_bytes, device=self.device)


_slots: dict = {}
_turn: dict = {}
_slots_lock = threading.Lock()      # feeder threads of several devices (analyze_images_batch(devices=...)) share this module


def _next_slot(device: torch.device) -> _Slot:
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    with _slots_lock:
        ring = _slots.setdefault(key, [_Slot(device), _Slot(device)])
        k = _turn.get(key, 0)
        _turn[key] = (k + 1) % len(ring)
        slot = ring[k]
    slot.lock.acquire()
    slot.wait()
    return slot


def release_buffers() -> None:
    """drops the cached pinned / device buffers (they are grow-only otherwise)"""
    with _slots_lock:
        for ring in _slots.values():
            for s in ring:
                s.wait()
        _slots.clear()
        _turn.clear()


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
        whs = np.zeros((max(self.n, 1), 3), dtype=np.int32)
        self.lib.iic_jpeg_plan_infos(h, whs.ctypes.data)
        self.whs = whs[:self.n]
        self.sizes: List[Tuple[int, int]] = [(int(r[1]), int(r[0])) for r in self.whs]       # (height, width); (0, 0) outside the envelope
        self.status: List[int] = [int(r[2]) for r in self.whs]
        self.reasons: List[str] = ["" if st == JPEG_OK else (self.lib.iic_jpeg_plan_reason(h, i) or b"").decode()
                                   for i, st in enumerate(self.status)]
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


def _decode(slot: _Slot, nbytes: int, offsets: np.ndarray) -> Tuple[List[Optional[torch.Tensor]], List[str]]:
    """slot.blob[:nbytes]: the files back to back in pinned memory."""
    device = slot.device
    plan = JpegPlan(slot.blob, offsets)
    try:
        out: List[Optional[torch.Tensor]] = [None] * plan.n
        ok = np.nonzero(plan.whs[:, 2] == JPEG_OK)[0]
        if ok.size == 0:
            return out, plan.reasons
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device)
            slot.reserve(nbytes, plan.staging_bytes, plan.scratch_bytes)
            slot.dev_blob[:nbytes + _PAD].copy_(slot.blob[:nbytes + _PAD], non_blocking=True)
            # one allocation for all decoded images of the batch (256-byte aligned slices)
            npix = plan.whs[ok, 0].astype(np.int64) * plan.whs[ok, 1].astype(np.int64) * 3
            sizes = (npix + 255) // 256 * 256
            starts = np.concatenate([[0], np.cumsum(sizes)[:-1]])
            pixels = torch.empty(int(sizes.sum()), dtype=torch.uint8, device=device)
            ptrs = np.zeros(plan.n, dtype=np.uint64)
            ptrs[ok] = np.uint64(pixels.data_ptr()) + starts.astype(np.uint64)
            for i, s0, nb in zip(ok.tolist(), starts.tolist(), npix.tolist()):
                out[i] = pixels[s0:s0 + nb].view(int(plan.whs[i, 1]), int(plan.whs[i, 0]), 3)
            rc = plan.lib.iic_jpeg_decode(plan.h, slot.dev_blob.data_ptr(), ptrs.ctypes.data_as(C.POINTER(C.c_void_p)),
                                          slot.staging.data_ptr(), slot.scratch.data_ptr(), stream.cuda_stream)
            if rc != L.IIC_OK:
                raise RuntimeError(f"iic_jpeg_decode failed (code {rc})")
            slot.event = torch.cuda.Event()
            slot.event.record(stream)
        return out, plan.reasons
    finally:
        plan.close()


def _device(device) -> torch.device:
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("the JPEG decoder runs on a CUDA device only (no CPU path)")
    return device


def decode_jpeg_bytes(files: Sequence[bytes], device) -> Tuple[List[Optional[torch.Tensor]], List[str]]:
    slot = _next_slot(_device(device))
    try:
        offsets = np.zeros(len(files) + 1, dtype=np.int64)
        np.cumsum([len(b) for b in files], out=offsets[1:])
        nbytes = int(offsets[-1])
        view = slot.host_blob(nbytes).numpy()
        for b, lo, hi in zip(files, offsets[:-1], offsets[1:]):
            view[lo:hi] = np.frombuffer(b, dtype=np.uint8)
        view[nbytes:nbytes + _PAD] = 0
        return _decode(slot, nbytes, offsets)
    finally:
        slot.lock.release()


def decode_jpeg_files(paths: Sequence[str], device) -> Tuple[List[Optional[torch.Tensor]], List[str]]:
    """Reads the files straight into ONE pinned buffer (a few reader threads, no per-file bytes objects) and decodes them on
    `device`.  An unreadable file gets (None, "<error>") like a file outside the envelope."""
    slot = _next_slot(_device(device))
    try:
        return _decode_files(slot, paths)
    finally:
        slot.lock.release()


def _decode_files(slot: _Slot, paths: Sequence[str]) -> Tuple[List[Optional[torch.Tensor]], List[str]]:
    sizes, errs = [], [""] * len(paths)
    for i, p in enumerate(paths):
        try:
            sizes.append(os.path.getsize(p))
        except OSError as e:
            sizes.append(0)
            errs[i] = f"unreadable: {e}"
    offsets = np.zeros(len(paths) + 1, dtype=np.int64)
    np.cumsum(sizes, out=offsets[1:])
    nbytes = int(offsets[-1])
    view = memoryview(slot.host_blob(nbytes).numpy())

    def read_some(first: int, step: int) -> None:
        # plain file descriptors (os.open / os.readv): the io module's wrappers cost more than the read of a 300 KB file
        for i in range(first, len(paths), step):
            if errs[i]:
                continue
            try:
                fd = os.open(paths[i], os.O_RDONLY)
                try:
                    got, want = 0, sizes[i]
                    while got < want:
                        n = os.readv(fd, [view[offsets[i] + got:offsets[i + 1]]])
                        if n <= 0:
                            break
                        got += n
                finally:
                    os.close(fd)
                if got != sizes[i]:
                    errs[i] = "short read"
            except OSError as e:
                errs[i] = f"unreadable: {e}"
            if errs[i] and sizes[i] >= 2:
                view[offsets[i]:offsets[i] + 2] = b"\0\0"       # no SOI: the plan reports the file as corrupt

    if len(paths) >= 4 * _READ_THREADS:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=_READ_THREADS) as ex:   # one strided share per thread; readv releases the GIL
            list(ex.map(lambda t: read_some(t, _READ_THREADS), range(_READ_THREADS)))
    else:
        read_some(0, 1)
    view[nbytes:nbytes + _PAD] = bytes(_PAD)
    imgs, reasons = _decode(slot, nbytes, offsets)
    return imgs, [e or r for e, r in zip(errs, reasons)]
