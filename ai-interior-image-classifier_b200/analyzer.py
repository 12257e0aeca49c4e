"""The reference's analyzer entry points, re-hosted on the CUDA engine.

Same class names, constructor arguments, method names and return shapes as /root/reference/main.py:
    InteriorImageDetector(device).is_interior_image(image, confidence_threshold) -> (bool, float, str)   main.py:149-226
    CachedInteriorAnalyzer(use_lora, lora_weights_path, lora_rank, lora_alpha, device)                     main.py:232-262
        .analyze_images_batch(image_paths, batch_size, filter_interiors, confidence_threshold) -> dict    main.py:371-469
        .analyze_image_from_url(url, filter_interiors)                                                     main.py:472-498
        ._analyze_image_tensor_fast(image_input)                                                           main.py:500-510
        .filter_interior_images(image_paths, confidence_threshold)                                         main.py:313-369
and the intended shape of python-worker/main_API.py's DatabaseStyleRoomAnalyzer._analyze_styles_batch (its body is
`pass` in the reference, main_API.py:268-271; the consumer at main_API.py:219-236 fixes the contract).

What changes underneath: one engine and ONE encode per image feeds both the detector head
and the five attribute heads - the reference encodes every interior image twice with two copies of the same
frozen ViT (main.py:333 and 444) - the 6 label groups (40 detector prompts + styles, characteristics, materials,
colors, room_types) are scored by one fused head kernel, and results come back in one device-to-host copy instead
of ~28 `.item()` syncs per image.  JSON label schema, prompt templates (main.py:302-305), top-5 per group, the
detector rule (main.py:216-220) and the result dict layout are unchanged.
"""
from __future__ import annotations

import json
import os
from concurrent.futures import ThreadPoolExecutor
from io import BytesIO
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from . import clip_compat as clip
from .lora import load_lora_weights_to_model, replace_linears_with_lora

DETECTOR_CATEGORIES = [
    # interiors (positive) - first 11, main.py:155-160, 185
    "interior of a room", "living room", "bedroom", "kitchen", "bathroom",
    "dining room", "office interior", "apartment interior", "house interior",
    "interior design", "home decor",
    # exteriors
    "building exterior", "outside of building", "street view", "garden",
    "landscape", "cityscape", "outdoor",
    # plans and diagrams
    "floor plan", "blueprint", "architectural plan", "diagram",
    "map", "technical drawing",
    # logos and graphics
    "company logo", "brand logo", "text", "signature",
    "advertisement", "brochure", "flyer",
    # other unwanted
    "person", "people", "animal", "pet", "car", "vehicle",
    "close-up of object", "product photo", "furniture close-up",
]
N_INTERIOR = 11
GROUP_ORDER = ("styles", "characteristics", "materials", "colors", "room_types")  # main.py:289-295


def load_image(path_or_url: str, timeout: int = 30):
    """main.py:119-128 / 325-327: URL -> requests, else local file; RGB PIL image or None."""
    from PIL import Image
    try:
        if path_or_url.startswith("http"):
            import requests
            r = requests.get(path_or_url, timeout=timeout)
            r.raise_for_status()
            return Image.open(BytesIO(r.content)).convert("RGB")
        return Image.open(path_or_url).convert("RGB")
    except Exception as e:  # noqa: BLE001 - the reference prints and skips
        print(f"Błąd ładowania {path_or_url}: {e}")
        return None


class DeviceImage:
    """An image that was decoded ON the GPU (jpeg.py): uint8 HWC tensor, ready for the preprocess kernel - no PIL object,
    no host copy of the pixels.  Quacks enough like a PIL image for the code that only forwards it (`.size`, `.convert`)."""

    def __init__(self, hwc_u8: torch.Tensor):
        self.tensor = hwc_u8
        self.size = (int(hwc_u8.shape[1]), int(hwc_u8.shape[0]))   # (W, H) like PIL

    def convert(self, mode: str):
        if mode != "RGB":
            raise ValueError("DeviceImage is always RGB")
        return self


GPU_DECODE_MIN_FILES = int(os.environ.get("IIC_GPU_DECODE_MIN_FILES", "4"))
INGEST_CHUNK = int(os.environ.get("IIC_INGEST_CHUNK", "1024"))


def load_images(paths: Sequence[str], device=None, gpu_decode: bool = False, timeout: int = 30):
    """Batch version of `load_image` (image ingest, main.py:322-346, 404-417).  With gpu_decode, the BYTES of
    local JPEG files are read into one pinned buffer and decoded on `device` by the engine's own decoder (csrc/jpeg.cu, jpeg.py:
    bit-identical to Pillow), so the pixels never exist on the host and the preprocess kernel reads them where they were decoded;
    everything else (URLs, PNG, progressive / CMYK JPEGs, an unreadable file) goes through `load_image` on a small thread pool,
    exactly like the reference."""
    out: List[object] = [None] * len(paths)
    rest = list(range(len(paths)))
    if gpu_decode and device is not None and torch.device(device).type == "cuda":
        from .jpeg import decode_jpeg_files
        idx = [i for i, p in enumerate(paths)
               if isinstance(p, str) and not p.startswith("http") and p.lower().endswith((".jpg", ".jpeg")) and os.path.isfile(p)]
        # Entropy decoding is serial inside a chain; the device decoder draws its parallelism from the files of the call and from
        # up to 8 chains inside each file.  Measured (1024x768, 308 KB files): 4 files 8.1 ms (Pillow on 4 threads: 12 ms), 16 files
        # 9.4 ms (40 ms), 64 files 7.7 k img/s, 1024 files 15 k img/s per blocking call.  One or two files are as fast on the host.
        if len(idx) >= GPU_DECODE_MIN_FILES:
            step = max(INGEST_CHUNK, 1) * 4                 # bounded staging / scratch however long the list is
            for lo in range(0, len(idx), step):
                part = idx[lo:lo + step]
                imgs, _ = decode_jpeg_files([paths[i] for i in part], device)
                for i, t in zip(part, imgs):
                    if t is not None:
                        out[i] = DeviceImage(t)
            rest = [i for i in rest if out[i] is None]
    if rest:
        with ThreadPoolExecutor(max_workers=4) as ex:   # host-side fetch/decode, as the reference does (main.py:345-346)
            for i, im in zip(rest, ex.map(lambda q: load_image(q, timeout), [paths[i] for i in rest])):
                out[i] = im
    return out


def iter_loaded(paths: Sequence[str], device=None, gpu_decode: bool = False, chunk: int = 0):
    """`load_images` over a long list, `chunk` paths at a time: yields (chunk_paths, images).  While the caller works on chunk i
    (preprocess + encoder + head), a worker thread already reads and parses the files of chunk i + 1 and enqueues their decode
    (file reads, header parsing and the ctypes calls release the GIL): the HOST side of the ingest disappears behind the encoder.
    The decode is enqueued on the caller's own stream - a side stream was measured and gains nothing: the encoder's GEMM CTAs
    hold the whole register file of an SM, so decode and encoder kernels cannot share one, and cross-stream allocation only
    makes the caching allocator's life hard.  Results do not depend on the chunking: every image is decoded and scored alone."""
    chunk = chunk or INGEST_CHUNK
    paths = list(paths)
    on_gpu = gpu_decode and device is not None and torch.device(device).type == "cuda"
    if len(paths) <= chunk or not on_gpu:
        if paths:
            yield paths, load_images(paths, device, gpu_decode)
        return
    dev = torch.device(device)
    stream = torch.cuda.current_stream(dev)

    def work(lo: int):
        part = paths[lo:lo + chunk]
        with torch.cuda.device(dev), torch.cuda.stream(stream):
            return part, load_images(part, dev, True)

    with ThreadPoolExecutor(max_workers=1) as ex:
        fut = ex.submit(work, 0)
        for lo in range(0, len(paths), chunk):
            part, imgs = fut.result()
            if lo + chunk < len(paths):
                fut = ex.submit(work, lo + chunk)
            yield part, imgs


def _encode_labels(model, texts: Sequence[str], device) -> torch.Tensor:
    with torch.no_grad():
        tok = clip.tokenize(list(texts)).to(device)
        f = model.encode_text(tok).float()
        return f / f.norm(dim=-1, keepdim=True)


class InteriorImageDetector:
    def __init__(self, device=None, model=None, preprocess=None):
        self.device = device or ("cuda" if torch.cuda.is_available() else "cpu")
        if model is None:
            model, preprocess = clip.load("ViT-B/16", device=self.device)
        self.model, self.preprocess = model, preprocess
        self.categories = list(DETECTOR_CATEGORIES)
        self.text_features = _encode_labels(self.model, self.categories, self.device)
        self.interior_indices = list(range(0, N_INTERIOR))
        self.non_interior_indices = list(range(N_INTERIOR, len(self.categories)))
        self._labels_key = None
        print(f"Detektor wnętrz zainicjalizowany. Kategorie wnętrz: {len(self.interior_indices)}, "
              f"inne: {len(self.non_interior_indices)}")

    def _engine(self):
        # the reference's detector is a separate, un-LoRA'd copy of the ViT (main.py:152 vs 241/247)
        eng = self.model.visual.sync_engine(use_lora=False)
        key = (id(eng), "detector")
        if getattr(eng, "_labels_owner", None) != key:
            eng.set_labels(self.text_features, [len(self.categories)], [N_INTERIOR], topk=1, logit_scale=100.0)
            eng._labels_owner = key
        return eng

    def detect_batch(self, images, confidence_threshold: float = 0.3, batch_size: int = 64):
        """List of PIL images -> [(is_interior, interior_confidence, top_category)], one encode per image, one
        device-to-host copy per batch.  Decision rule: main.py:208-222."""
        out = []
        for i in range(0, len(images), batch_size):
            # sync (weights / LoRA / labels) + classify as ONE critical section of the engine: no other thread's upload can
            # interleave (main.py:345-346 calls the detector from a 4-thread pool)
            with self.model.visual.engine()._lock:
                eng = self._engine()
                u8 = [clip.Preprocess._to_u8(im, eng.device) for im in images[i:i + batch_size]]
                r = eng.classify(u8, want_embedding=False)
                top_conf, top_idx = r.topk_val[:, 0, 0].cpu(), r.topk_idx[:, 0, 0].cpu()
                interior = r.split_sum[:, 0].cpu()
                non_interior = r.probs[:, N_INTERIOR:].sum(-1).cpu()
            for j in range(len(u8)):
                ok = bool(interior[j] > non_interior[j] and top_conf[j] > confidence_threshold)
                out.append((ok, float(interior[j]), self.categories[int(top_idx[j])]))
        return out

    def is_interior_image(self, image, confidence_threshold: float = 0.3):
        if image is None:
            return False, 0.0, "invalid image"
        try:
            return self.detect_batch([image], confidence_threshold)[0]
        except Exception as e:  # noqa: BLE001 - reference behaviour, main.py:224-226
            print(f"Błąd podczas detekcji wnętrza: {e}")
            return False, 0.0, f"error: {str(e)}"


class CachedInteriorAnalyzer:
    def __init__(self, use_lora: bool = False, lora_weights_path: Optional[str] = None, lora_rank: int = 4,
                 lora_alpha: float = 8, device=None, json_path: str = "interior_dataset.json", model=None,
                 preprocess=None, share_detector_encoder: Optional[bool] = None):
        self.device = device or ("cuda" if torch.cuda.is_available() else "cpu")
        print(f"Urządzenie: {self.device}")
        if model is None:
            model, preprocess = clip.load("ViT-B/16", device=self.device)
        self.model, self.preprocess = model, preprocess
        # The reference gives the detector its own, un-LoRA'd ViT (main.py:238).  Text features of the detector must
        # therefore come from the BASE text tower: encode them before LoRA is applied to the shared model.
        self.detector = InteriorImageDetector(device=self.device, model=self.model, preprocess=self.preprocess)
        self.use_lora = False
        if use_lora:
            print("Aplikuję LoRA do modelu...")
            replaced = replace_linears_with_lora(self.model, rank=lora_rank, alpha=lora_alpha)
            print(f"Zastąpiono warstwy Linear: {len(replaced)}")
            if lora_weights_path and os.path.exists(lora_weights_path):
                print("Wczytywanie wag LoRA...")
                load_lora_weights_to_model(self.model, lora_weights_path, strict_match=False)
            else:
                print("Brak ścieżki do wag LoRA -> używam losowych LoRA")
            self.use_lora = True
        else:
            print("Nie używam LoRA - model bez modyfikacji")
        # One encode can serve both heads only while the LoRA'd vision tower equals the base tower (always true for
        # the shipped checkpoints: they hold text-tower tensors only and lora_B initialises to zero - SURVEY F7).
        # local JPEG files are decoded on the GPU (csrc/jpeg.cu: bit-identical to Pillow, so the results do not
        # depend on the switch); IIC_GPU_DECODE=0 keeps every file on the reference's host pool.  See load_images().
        self.gpu_decode = os.environ.get("IIC_GPU_DECODE", "1") == "1"
        self._vision_lora_is_zero = self._check_vision_lora_zero()
        self.share_detector_encoder = (self._vision_lora_is_zero if share_detector_encoder is None
                                       else share_detector_encoder)
        self.training_data = self._load_training_data(json_path)
        self.all_categories = self._extract_all_categories()
        self.text_features_cache: Dict[str, torch.Tensor] = {}
        self._precompute_text_features_optimized()

    # -- reference-shaped helpers ---------------------------------------------------------------------------
    def _check_vision_lora_zero(self) -> bool:
        for n, p in self.model.visual.named_parameters():
            if n.endswith("lora.lora_B") and ".attn.out_proj." not in n and bool((p != 0).any()):
                return False
        return True

    def _load_training_data(self, json_path: str = "interior_dataset.json"):
        try:
            with open(json_path, "r", encoding="utf-8") as f:
                return json.load(f).get("training_data", [])
        except Exception as e:  # noqa: BLE001
            print(f"Nie udało się wczytać training data: {e}")
            return []

    def _extract_all_categories(self) -> Dict[str, List[str]]:
        """interior_dataset.json schema (main.py:273-295).  Label order inside a group is sorted here; the
        reference's order comes from set iteration and is not stable between runs - results are keyed
        by label string either way."""
        groups = {k: set() for k in GROUP_ORDER}
        for item in self.training_data:
            groups["styles"].add(item.get("style", ""))
            groups["room_types"].add(item.get("room_type", ""))
            for key in ("characteristics", "materials", "colors"):
                groups[key].update(item.get(key, []))
        return {k: sorted(v for v in groups[k] if v) for k in GROUP_ORDER}

    def _precompute_text_features_optimized(self):
        print("Prekomputowanie cech tekstowych...")
        for category, attributes in self.all_categories.items():
            if not attributes:
                continue
            texts = [f"{a}" for a in attributes] if category == "room_types" else [f"wnętrze z {a}" for a in attributes]
            self.text_features_cache[category] = _encode_labels(self.model, texts, self.device)
        print("Prekomputowanie zakończone.")

    # -- fused scoring ---------------------------------------------------------------------------------------
    def _group_names(self) -> List[str]:
        return [g for g in GROUP_ORDER if g in self.text_features_cache]

    def _engine(self, with_detector: bool):
        eng = self.model.visual.sync_engine()
        names = self._group_names()
        key = (id(eng), "analyzer", with_detector, tuple(names))
        if getattr(eng, "_labels_owner", None) != key:
            mats = [self.text_features_cache[g] for g in names]
            sizes = [m.shape[0] for m in mats]
            split = [0] * len(names)
            if with_detector:
                mats = [self.detector.text_features] + mats
                sizes = [len(self.detector.categories)] + sizes
                split = [N_INTERIOR] + split
            eng.set_labels(torch.cat(mats, 0), sizes, split, topk=5, logit_scale=100.0)
            eng._labels_owner = key
        return eng, names

    def _analysis_from_head(self, tv, ti, row: int, names: List[str], g0: int) -> Dict[str, list]:
        out = {}
        for gi, g in enumerate(names):
            attrs = self.all_categories[g]
            k = min(5, len(attrs))
            out[g] = [(attrs[int(ti[row, g0 + gi, j])], float(tv[row, g0 + gi, j])) for j in range(k)]
        return out

    def _classify_pil(self, images, with_detector: bool, batch_size: int):
        """PIL list -> (topk_val, topk_idx, split_sum, probs-of-detector) on the host, one D2H per batch."""
        tvs, tis, sss, nons = [], [], [], []
        for i in range(0, len(images), batch_size):
            with self.model.visual.engine()._lock:      # sync + labels + classify: one critical section
                eng, names = self._engine(with_detector)
                u8 = [clip.Preprocess._to_u8(im, eng.device) for im in images[i:i + batch_size]]
                r = eng.classify(u8, want_embedding=False)
                tvs.append(r.topk_val.cpu())
                tis.append(r.topk_idx.cpu())
                sss.append(r.split_sum.cpu())
                if with_detector:
                    nons.append(r.probs[:, N_INTERIOR:len(self.detector.categories)].sum(-1).cpu())
        cat = lambda xs: torch.cat(xs, 0) if xs else None
        return cat(tvs), cat(tis), cat(sss), cat(nons), names

    # -- public API (reference signatures) -------------------------------------------------------------------
    def filter_interior_images(self, image_paths, confidence_threshold: float = 0.3):
        print(f" Filtrowanie {len(image_paths)} obrazów - wykrywanie wnętrz...")
        imgs = load_images(image_paths, self.device, self.gpu_decode)   # fetch/decode only; GPU work is batched below
        ok = [(p, im) for p, im in zip(image_paths, imgs) if im is not None]
        interior_images, non_interior_info = [], []
        for p, im in zip(image_paths, imgs):
            if im is None:
                non_interior_info.append({"path": p, "confidence": 0.0, "category": "load error",
                                          "reason": "Nie wnętrze: load error (confidence: 0.000)"})
        dets = self.detector.detect_batch([im for _, im in ok], confidence_threshold) if ok else []
        for (p, im), (is_interior, confidence, category) in zip(ok, dets):
            if is_interior:
                interior_images.append((p, im, confidence))
            else:
                non_interior_info.append({"path": p, "confidence": confidence, "category": category,
                                          "reason": f"Nie wnętrze: {category} (confidence: {confidence:.3f})"})
        print(f" Znaleziono {len(interior_images)} obrazów wnętrz")
        print(f" Odrzucono {len(non_interior_info)} obrazów nie-wnętrz")
        return interior_images, non_interior_info

    def _analyze_shared(self, image_paths, batch_size: int, confidence_threshold: float):
        """filter + analyse in ONE encode per image (valid while the vision LoRA delta is zero)."""
        results = {}
        for part, imgs in iter_loaded(image_paths, self.device, self.gpu_decode):
            self._analyze_shared_chunk(part, imgs, results, batch_size, confidence_threshold)
        return {p: results[p] for p in image_paths if p in results}      # the reference's order: that of image_paths

    def _analyze_shared_chunk(self, image_paths, imgs, results, batch_size: int, confidence_threshold: float):
        ok = [(p, im) for p, im in zip(image_paths, imgs) if im is not None]
        for p, im in zip(image_paths, imgs):
            if im is None:
                results[p] = {"is_interior": False, "interior_confidence": 0.0, "detected_category": "load error",
                              "analysis": {}, "reason": "Nie wnętrze: load error (confidence: 0.000)"}
        if not ok:
            return results
        # GPU batch: results do not depend on it; long lists of device-decoded images run at the encoder's efficient batch sizes
        gpu_batch = max(batch_size, 256 if len(ok) >= 256 else 64)
        tv, ti, ss, non, names = self._classify_pil([im for _, im in ok], True, gpu_batch)
        for r, (p, _) in enumerate(ok):
            interior, top_conf = float(ss[r, 0]), float(tv[r, 0, 0])
            category = self.detector.categories[int(ti[r, 0, 0])]
            if interior > float(non[r]) and top_conf > confidence_threshold:
                results[p] = {"is_interior": True, "interior_confidence": interior, "detected_category": "interior",
                              "analysis": self._analysis_from_head(tv, ti, r, names, 1),
                              "reason": "Success - interior image analyzed"}
            else:
                results[p] = {"is_interior": False, "interior_confidence": interior, "detected_category": category,
                              "analysis": {}, "reason": f"Nie wnętrze: {category} (confidence: {interior:.3f})"}
        return results

    # -- data parallel inside one process (images are independent, weights replicated, no collective) ------------
    @staticmethod
    def _cuda_device(device) -> torch.device:
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if dev.type != "cuda":
            raise RuntimeError(f"analyze_images_batch(devices=...): {dev} is not a CUDA device (there is no CPU path)")
        return dev if dev.index is not None else torch.device("cuda", torch.cuda.current_device())

    def _make_replica(self, dev: torch.device) -> "CachedInteriorAnalyzer":
        """This analyzer bound to `dev`: a deep copy of the model (hence its own engine and workspace) with the cached label
        matrices moved over; label schema, categories and settings are shared."""
        import copy
        r = copy.copy(self)                       # shallow: shares training_data / all_categories
        r.__dict__.pop("_replicas", None)
        r.device = str(dev)
        r.model = copy.deepcopy(self.model).to(dev)
        r.model.visual._engine, r.model.visual._sig, r.model.visual._zero_b = None, None, {}   # a fresh engine on `dev`
        if hasattr(r.model, "_text_engine"):
            r.model._text_engine, r.model._text_sig = None, None
        r.preprocess = clip.Preprocess(r.model.visual)
        r.detector = copy.copy(self.detector)
        r.detector.device, r.detector.model, r.detector.preprocess = str(dev), r.model, r.preprocess
        r.detector.text_features = self.detector.text_features.to(dev)
        r.text_features_cache = {g: t.to(dev) for g, t in self.text_features_cache.items()}
        return r

    def _analyze_data_parallel(self, image_paths, devices, **kw):
        """Contiguous shards of `image_paths`, one feeder thread per entry of `devices`; every thread runs the ordinary
        single-device entry point on its replica (ctypes releases the GIL, so the engines run concurrently).  The merged dict
        equals the single-device result: an image's result does not depend on the batch it travels in.  A device may be named
        more than once (several engines on one GPU)."""
        from .dp import shard_bounds
        paths = list(image_paths)
        devs = [self._cuda_device(d) for d in devices]
        cache = self.__dict__.setdefault("_replicas", {})
        own = self._cuda_device(self.device)
        reps = []
        for slot, dev in enumerate(devs):
            if dev == own and self not in reps:
                reps.append(self)
                continue
            key = (slot, str(dev))
            if key not in cache:
                cache[key] = self._make_replica(dev)
            reps.append(cache[key])
        spans = [shard_bounds(len(paths), k, len(reps)) for k in range(len(reps))]

        def work(k):
            lo, hi = spans[k]
            if lo == hi:
                return {}
            with torch.cuda.device(torch.device(reps[k].device)):
                return reps[k].analyze_images_batch(paths[lo:hi], **kw)
        with ThreadPoolExecutor(max_workers=len(reps)) as ex:
            parts = list(ex.map(work, range(len(reps))))
        merged = {}
        for part in parts:
            merged.update(part)
        return merged

    def analyze_images_batch(self, image_paths, batch_size: int = 16, filter_interiors: bool = True,
                             confidence_threshold: float = 0.3, devices: Optional[Sequence] = None):
        """main.py:371-469.  `devices` (new, optional): a list of CUDA devices, e.g. range(8) - the paths are sharded over them
        (one engine + one feeder thread per device); the returned dict is the same as on one device."""
        if devices is not None and len(list(devices)) > 1:
            return self._analyze_data_parallel(image_paths, list(devices), batch_size=batch_size, filter_interiors=filter_interiors,
                                               confidence_threshold=confidence_threshold)
        results, valid_images, image_metadata = {}, [], []
        if filter_interiors and self.share_detector_encoder:
            return self._analyze_shared(image_paths, batch_size, confidence_threshold)
        if filter_interiors:
            interior_images, non_interior_info = self.filter_interior_images(image_paths, confidence_threshold)
            for info in non_interior_info:
                results[info["path"]] = {"is_interior": False, "interior_confidence": info["confidence"],
                                         "detected_category": info["category"], "analysis": {}, "reason": info["reason"]}
            for path, img, confidence in interior_images:
                valid_images.append(img)
                image_metadata.append({"path": path, "interior_confidence": confidence, "is_interior": True})
        else:
            print("  Pomijam filtrowanie wnętrz - przetwarzam wszystkie obrazy")
            if len(image_paths) > INGEST_CHUNK and self.gpu_decode:      # long list: ingest of chunk i + 1 under the encode of chunk i
                for part, imgs in iter_loaded(image_paths, self.device, self.gpu_decode):
                    results.update(self._analyze_no_filter_chunk(part, imgs, batch_size))
                return {p: results[p] for p in image_paths if p in results}
            for path, img in zip(image_paths, load_images(image_paths, self.device, self.gpu_decode)):
                if img is not None:
                    valid_images.append(img)
                    image_metadata.append({"path": path, "interior_confidence": 1.0, "is_interior": True})
                else:
                    results[path] = {"is_interior": False, "interior_confidence": 0.0, "detected_category": "load error",
                                     "analysis": {}, "reason": "Błąd ładowania"}
        if not valid_images:
            print("Brak obrazów do analizy")
            return results
        print(f"  Przetwarzam {len(valid_images)} obrazów w batchach po {batch_size}...")
        tv, ti, _, _, names = self._classify_pil(valid_images, False, max(batch_size, 1))
        for idx, meta in enumerate(image_metadata):
            results[meta["path"]] = {"is_interior": True, "interior_confidence": meta["interior_confidence"],
                                     "detected_category": "interior",
                                     "analysis": self._analysis_from_head(tv, ti, idx, names, 0),
                                     "reason": "Success - interior image analyzed"}
        return results

    def _analyze_no_filter_chunk(self, image_paths, imgs, batch_size: int):
        """the filter_interiors=False body of analyze_images_batch (main.py:404-467) for one chunk of a long list"""
        out = {}
        ok = [(p, im) for p, im in zip(image_paths, imgs) if im is not None]
        for p, im in zip(image_paths, imgs):
            if im is None:
                out[p] = {"is_interior": False, "interior_confidence": 0.0, "detected_category": "load error",
                          "analysis": {}, "reason": "Błąd ładowania"}
        if ok:
            tv, ti, _, _, names = self._classify_pil([im for _, im in ok], False, max(batch_size, 256 if len(ok) >= 256 else 1))
            for idx, (p, _) in enumerate(ok):
                out[p] = {"is_interior": True, "interior_confidence": 1.0, "detected_category": "interior",
                          "analysis": self._analysis_from_head(tv, ti, idx, names, 0), "reason": "Success - interior image analyzed"}
        return out

    def analyze_image_from_url(self, url, filter_interiors: bool = True):
        img = load_image(url)
        if img is None:
            return {"is_interior": False, "reason": "Failed to load image"}
        confidence = 1.0
        if filter_interiors:
            is_interior, confidence, category = self.detector.is_interior_image(img)
            if not is_interior:
                return {"is_interior": False, "interior_confidence": confidence, "detected_category": category,
                        "analysis": {}, "reason": f"Not an interior image: {category}"}
        tv, ti, _, _, names = self._classify_pil([img], False, 1)
        return {"is_interior": True, "interior_confidence": confidence if filter_interiors else 1.0,
                "detected_category": "interior", "analysis": self._analysis_from_head(tv, ti, 0, names, 0),
                "reason": "Success - interior image analyzed"}

    def _analyze_image_tensor_fast(self, image_input: torch.Tensor):
        """image_input: preprocessed float tensor [1,3,R,R] (what the reference passes, main.py:489, 500)."""
        with self.model.visual.engine()._lock:
            eng, names = self._engine(False)
            r = eng.classify_patches(eng.patchify(image_input.to(eng.device)), image_input.shape[0], want_embedding=False)
            tv, ti = r.topk_val.cpu(), r.topk_idx.cpu()
        return self._analysis_from_head(tv, ti, 0, names, 0)


class DatabaseStyleRoomAnalyzer:
    """Compute half of python-worker/main_API.py:130-281 (the Mongo plumbing stays where it is): ten fixed Polish
    style prompts `wnętrze w stylu {style}` -> per image {'style': argmax label, 'confidence': max prob}."""

    STYLES = ["nowoczesny", "tradycyjny", "skandynawski", "industrialny", "minimalistyczny", "rustykalny", "glamour",
              "boho", "klasyczny", "loft"]

    def __init__(self, use_lora: bool = False, lora_weights_path: Optional[str] = None, device=None, styles=None,
                 model=None, preprocess=None):
        self.device = device or ("cuda" if torch.cuda.is_available() else "cpu")
        if model is None:
            model, preprocess = clip.load("ViT-B/16", device=self.device)
        self.model, self.preprocess = model, preprocess
        self.detector = InteriorImageDetector(device=self.device, model=self.model, preprocess=self.preprocess)
        if use_lora:
            replace_linears_with_lora(self.model, rank=4, alpha=8)  # main_API.py:143
            if lora_weights_path and os.path.exists(lora_weights_path):
                load_lora_weights_to_model(self.model, lora_weights_path, strict_match=False)
        self.styles = list(styles) if styles is not None else list(self.STYLES)
        self.style_features = _encode_labels(self.model, [f"wnętrze w stylu {s}" for s in self.styles], self.device)

    def _analyze_styles_batch(self, images, batch_size: int = 16):
        out = []
        for i in range(0, len(images), batch_size):
            with self.model.visual.engine()._lock:
                eng = self.model.visual.sync_engine()
                eng.set_labels(self.style_features, [len(self.styles)], None, topk=1, logit_scale=100.0)
                eng._labels_owner = None
                u8 = [clip.Preprocess._to_u8(im, eng.device) for im in images[i:i + batch_size]]
                r = eng.classify(u8, want_embedding=False)
                tv, ti = r.topk_val.cpu(), r.topk_idx.cpu()
            out += [{"style": self.styles[int(ti[j, 0, 0])], "confidence": float(tv[j, 0, 0])} for j in range(tv.shape[0])]
        return out
