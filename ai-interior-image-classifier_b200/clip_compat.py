"""Drop-in for what the reference takes from the `clip` package: `load`, `tokenize`, `available_models`.

    model, preprocess = load("ViT-B/16", device="cuda")        # /root/reference/main.py:152, 241; train_lora.py:174
    model.encode_image(batch)   -> CUDA engine (csrc/, through the C ABI)      main.py:204, 444, 503
    model.encode_text(tokens)   -> PyTorch by default (label embeddings are an INPUT of the hot path);
                                   with `model.text_on_engine = True` the same CUDA engine runs the text tower
                                   (causal attention, live text LoRA: SURVEY.md 8(f) row N3)
    preprocess(PIL.Image)       -> CUDA preprocess kernel, returns Tensor[3,R,R] like clip._transform

The module tree and parameter names are OpenAI CLIP's (`visual.transformer.resblocks.{i}.mlp.c_fc.weight`, ...), so
the reference's name-based LoRA code - `replace_linears_with_lora` (main.py:62-74), `load_lora_weights_to_model`
(main.py:86-113), `LoRACLIPWrapper` (train_lora.py:47-100) - works on this model unchanged, and shipped
`lora_models/*.pth` checkpoints load by the same suffix matching.

The vision tower has NO PyTorch forward.  `encode_image` collects the current parameters (including any LoRA pair
hanging off `mlp.c_fc` / `mlp.c_proj`), hands them to the engine and runs the sm_100a kernels; on a machine without
a B200 or without the built extension it raises.
"""
from __future__ import annotations

import math
import os
import threading
import warnings
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from .engine import Engine, VisionArch

_MODELS = {
    "ViT-B/16": dict(embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768, vision_patch_size=16,
                     context_length=77, vocab_size=49408, transformer_width=512, transformer_heads=8,
                     transformer_layers=12, file="ViT-B-16.pt"),
    "ViT-L/14@336px": dict(embed_dim=768, image_resolution=336, vision_layers=24, vision_width=1024,
                           vision_patch_size=14, context_length=77, vocab_size=49408, transformer_width=768,
                           transformer_heads=12, transformer_layers=12, file="ViT-L-14-336px.pt"),
}
_SOT, _EOT = 49406, 49407
_ENGINE_CREATE_LOCK = threading.Lock()


def _opted_in(flag: Optional[bool], env: str) -> bool:
    return bool(flag) if flag is not None else os.environ.get(env, "0") == "1"


def available_models() -> List[str]:
    return list(_MODELS)


class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class _Fp32LayerNorm(nn.LayerNorm):
    def forward(self, x):
        return F.layer_norm(x.float(), self.normalized_shape, self.weight, self.bias, self.eps).to(x.dtype)


class ResidualAttentionBlock(nn.Module):
    """Parameter container with upstream names; `forward` is only used by the TEXT tower."""

    def __init__(self, d_model: int, n_head: int, causal: bool):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = _Fp32LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict(c_fc=nn.Linear(d_model, 4 * d_model), gelu=QuickGELU(),
                                             c_proj=nn.Linear(4 * d_model, d_model)))
        self.ln_2 = _Fp32LayerNorm(d_model)
        self.causal = causal

    def forward(self, x):  # x: [T, N, d] (sequence first, as upstream)
        y = self.ln_1(x)
        mask = None
        if self.causal:
            T = x.shape[0]
            mask = torch.full((T, T), float("-inf"), device=x.device, dtype=x.dtype).triu_(1)
        x = x + self.attn(y, y, y, need_weights=False, attn_mask=mask)[0]
        return x + self.mlp(self.ln_2(x))


class Transformer(nn.Module):
    def __init__(self, width: int, layers: int, heads: int, causal: bool):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, causal) for _ in range(layers)])

    def forward(self, x):
        return self.resblocks(x)


def _is_lora_wrapped(m: nn.Module) -> bool:
    """Duck type of the reference's LoRALinear (main.py:34-58, train_lora.py:32-44) and of our mirror."""
    lo = getattr(m, "lora", None)
    return hasattr(m, "linear") and lo is not None and hasattr(lo, "lora_A") and hasattr(lo, "lora_B")


class VisionTransformer(nn.Module):
    """Holds the `visual.*` parameters; computes through the CUDA engine."""

    def __init__(self, input_resolution: int, patch_size: int, width: int, layers: int, heads: int, output_dim: int):
        super().__init__()
        self.input_resolution, self.patch_size, self.output_dim = input_resolution, patch_size, output_dim
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        s = width ** -0.5
        self.class_embedding = nn.Parameter(s * torch.randn(width))
        self.positional_embedding = nn.Parameter(s * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = _Fp32LayerNorm(width)
        self.transformer = Transformer(width, layers, heads, causal=False)
        self.ln_post = _Fp32LayerNorm(width)
        self.proj = nn.Parameter(s * torch.randn(width, output_dim))
        self.arch = VisionArch(input_resolution, patch_size, width, layers, heads, output_dim, L.ACT_QUICK_GELU)
        # Reference parity: nn.MultiheadAttention never calls out_proj(x), so a LoRA hanging off
        # attn.out_proj has no effect in the reference.  Set True to apply it anyway (generic slot of the kernel).
        self.apply_out_proj_lora = False
        # 16-bit operand format of the engine: "f16" (the default, _lib.DEFAULT_OPERAND_DTYPE: upstream CLIP's own GPU
        # dtype) or "bf16" (explicit non-default arm).
        self.operand_dtype = L.operand_dtype_name(None)
        self._engine: Optional[Engine] = None
        self._sig = None
        self._zero_b: Dict[int, Tuple[int, int, bool]] = {}   # id(lora_B) -> (data_ptr, _version, "is all zero")

    # -- engine plumbing ------------------------------------------------------------------------------------
    def engine(self) -> Engine:
        dev = self.conv1.weight.device
        if dev.type != "cuda":
            raise RuntimeError(
                "the iic-b200 vision tower only runs on a CUDA device (B200); move the model with .to('cuda'). "
                "There is deliberately no CPU fallback.")
        want = torch.float16 if L.operand_dtype_name(self.operand_dtype) == "f16" else torch.bfloat16
        with _ENGINE_CREATE_LOCK:
            if self._engine is None or self._engine.device != dev or self._engine.op_dtype != want:
                self._engine = Engine(self.arch, dev, operand_dtype=want)
                self._sig = None
            return self._engine

    def _b_is_zero(self, b: torch.Tensor) -> bool:
        """lora_B == 0 (fresh wrap / tensor missing from the checkpoint: the delta is exactly zero).  The answer
        needs a device-to-host sync, so it is cached per tensor on (data_ptr, _version): steady-state calls make none."""
        key = id(b)
        hit = self._zero_b.get(key)
        if hit is not None and hit[0] == b.data_ptr() and hit[1] == b._version:
            return hit[2]
        z = not bool((b != 0).any())
        self._zero_b[key] = (b.data_ptr(), b._version, z)
        return z

    def _tensors(self, keep_zero_lora: bool = False) -> Tuple[Dict[str, torch.Tensor], Dict[Tuple[int, int], Tuple[torch.Tensor, torch.Tensor, float]]]:
        sd: Dict[str, torch.Tensor] = {
            "conv1.weight": self.conv1.weight, "class_embedding": self.class_embedding,
            "positional_embedding": self.positional_embedding, "ln_pre.weight": self.ln_pre.weight,
            "ln_pre.bias": self.ln_pre.bias, "ln_post.weight": self.ln_post.weight, "ln_post.bias": self.ln_post.bias,
            "proj": self.proj,
        }
        lora: Dict[Tuple[int, int], Tuple[torch.Tensor, torch.Tensor, float]] = {}
        for i, blk in enumerate(self.transformer.resblocks):
            p = f"transformer.resblocks.{i}."
            sd[p + "ln_1.weight"], sd[p + "ln_1.bias"] = blk.ln_1.weight, blk.ln_1.bias
            sd[p + "ln_2.weight"], sd[p + "ln_2.bias"] = blk.ln_2.weight, blk.ln_2.bias
            sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"] = blk.attn.in_proj_weight, blk.attn.in_proj_bias
            for which, name, mod in ((L.LORA_OUT_PROJ, "attn.out_proj", blk.attn.out_proj),
                                     (L.LORA_C_FC, "mlp.c_fc", blk.mlp.c_fc), (L.LORA_C_PROJ, "mlp.c_proj", blk.mlp.c_proj)):
                sd[p + name + ".weight"], sd[p + name + ".bias"] = mod.weight, mod.bias  # proxies on a LoRALinear
                if _is_lora_wrapped(mod) and (which != L.LORA_OUT_PROJ or self.apply_out_proj_lora):
                    scaling = float(getattr(mod.lora, "scaling", 1.0))
                    if keep_zero_lora or not self._b_is_zero(mod.lora.lora_B):  # B == 0 (fresh / missing in ckpt): delta is exactly 0
                        lora[(i, which)] = (mod.lora.lora_A, mod.lora.lora_B, scaling)
        return sd, lora

    def sync_engine(self, force: bool = False, use_lora: bool = True, keep_zero_lora: bool = False) -> Engine:
        """(Re)upload whatever changed since the last call (optimizer step, checkpoint load, `.data` swap).
        use_lora=False runs the frozen base tower (the reference's detector owns an un-LoRA'd copy, main.py:238)."""
        eng = self.engine()
        # the whole comparison + upload holds the engine lock (re-entrant): a concurrent classify on another thread can never
        # see a half-switched LoRA configuration (the reference calls its detector from a 4-thread pool, main.py:345-346)
        with eng._lock:
            return self._sync_locked(eng, force, use_lora, keep_zero_lora)

    def _sync_locked(self, eng: Engine, force: bool, use_lora: bool, keep_zero_lora: bool) -> Engine:
        sd, lora = self._tensors(keep_zero_lora)
        if not use_lora:
            lora = {}
        sig_w = tuple((k, t.data_ptr(), t._version, t.device.type) for k, t in sd.items())
        sig_l = tuple((k, a.data_ptr(), a._version, b.data_ptr(), b._version, s) for k, (a, b, s) in sorted(lora.items()))
        old_w, old_l = self._sig if self._sig is not None else (None, None)
        with torch.no_grad():
            if force or sig_w != old_w:
                eng.load_visual_state_dict(sd)
            if force or sig_l != old_l:
                for i in range(self.arch.layers):
                    for which in (L.LORA_IN_PROJ, L.LORA_OUT_PROJ, L.LORA_C_FC, L.LORA_C_PROJ):
                        if (i, which) in lora:
                            a, b, s = lora[(i, which)]
                            eng.set_lora(i, which, a, b, s)
                        else:
                            eng.set_lora(i, which, None, None)
        self._sig = (sig_w, sig_l)
        return eng

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            pass  # inference engine: gradients do not flow through this call (see train.py for the training step)
        eng = self.sync_engine()
        return eng.encode_image(x.to(eng.device)).to(x.dtype if x.is_floating_point() else torch.float32)


class CLIP(nn.Module):
    def __init__(self, embed_dim, image_resolution, vision_layers, vision_width, vision_patch_size, context_length,
                 vocab_size, transformer_width, transformer_heads, transformer_layers, **_unused):
        super().__init__()
        self.context_length = context_length
        self.vocab_size = vocab_size
        self.visual = VisionTransformer(image_resolution, vision_patch_size, vision_width, vision_layers,
                                        vision_width // 64, embed_dim)
        self.transformer = Transformer(transformer_width, transformer_layers, transformer_heads, causal=True)
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(context_length, transformer_width))
        self.ln_final = _Fp32LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))
        # run encode_text on the CUDA engine too (same GEMM / LayerNorm / attention kernels, causal mask,
        # T = 77, LoRA pairs hanging off the text MLPs fused exactly like the vision ones).  Off by default: the label
        # matrix is an input of the image hot path and the reference computes it in fp32.
        self.text_on_engine = os.environ.get("IIC_TEXT_ON_ENGINE", "0") == "1"
        self._text_engine: Optional[Engine] = None
        self._text_sig = None
        self._text_arch = VisionArch(image_size=224, patch_size=16, width=transformer_width, layers=transformer_layers,
                                     heads=transformer_heads, embed_dim=embed_dim, activation=L.ACT_QUICK_GELU,
                                     seq_tokens=context_length, causal=True)
        self._init()

    def _init(self):
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        for t in (self.transformer, self.visual.transformer):
            attn_std, fc_std = t.width ** -0.5, (2 * t.width) ** -0.5
            proj_std = attn_std * (2 * t.layers) ** -0.5
            for b in t.resblocks:
                nn.init.normal_(b.attn.in_proj_weight, std=attn_std)
                nn.init.normal_(b.attn.out_proj.weight, std=proj_std)
                nn.init.normal_(b.mlp.c_fc.weight, std=fc_std)
                nn.init.normal_(b.mlp.c_proj.weight, std=proj_std)
        nn.init.normal_(self.text_projection, std=self.transformer.width ** -0.5)

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    def encode_image(self, image: torch.Tensor) -> torch.Tensor:
        return self.visual(image)

    # -- text tower on the engine ---------------------------------------------------------------------------------
    def sync_text_engine(self, force: bool = False) -> Engine:
        dev = self.token_embedding.weight.device
        if dev.type != "cuda":
            raise RuntimeError("text_on_engine needs the model on a CUDA device (B200); there is no CPU fallback for the engine")
        v = self.visual
        want = torch.float16 if L.operand_dtype_name(v.operand_dtype) == "f16" else torch.bfloat16
        with _ENGINE_CREATE_LOCK:
            if self._text_engine is None or self._text_engine.device != dev or self._text_engine.op_dtype != want:
                self._text_engine = Engine(self._text_arch, dev, operand_dtype=want)
                self._text_sig = None
            eng = self._text_engine
        with eng._lock:
            return self._sync_text_locked(eng, force)

    def _sync_text_locked(self, eng: Engine, force: bool) -> Engine:
        v = self.visual
        sd: Dict[str, torch.Tensor] = {"ln_final.weight": self.ln_final.weight, "ln_final.bias": self.ln_final.bias,
                                       "text_projection": self.text_projection}
        lora: Dict[Tuple[int, int], Tuple[torch.Tensor, torch.Tensor, float]] = {}
        for i, blk in enumerate(self.transformer.resblocks):
            p = f"transformer.resblocks.{i}."
            sd[p + "ln_1.weight"], sd[p + "ln_1.bias"] = blk.ln_1.weight, blk.ln_1.bias
            sd[p + "ln_2.weight"], sd[p + "ln_2.bias"] = blk.ln_2.weight, blk.ln_2.bias
            sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"] = blk.attn.in_proj_weight, blk.attn.in_proj_bias
            for which, name, mod in ((L.LORA_OUT_PROJ, "attn.out_proj", blk.attn.out_proj),
                                     (L.LORA_C_FC, "mlp.c_fc", blk.mlp.c_fc), (L.LORA_C_PROJ, "mlp.c_proj", blk.mlp.c_proj)):
                sd[p + name + ".weight"], sd[p + name + ".bias"] = mod.weight, mod.bias  # proxies on a LoRALinear
                # an attn.out_proj LoRA is dead in the reference's forward (F4): never applied
                if _is_lora_wrapped(mod) and which != L.LORA_OUT_PROJ and not v._b_is_zero(mod.lora.lora_B):
                    lora[(i, which)] = (mod.lora.lora_A, mod.lora.lora_B, float(getattr(mod.lora, "scaling", 1.0)))
        sig_w = tuple((k, t.data_ptr(), t._version) for k, t in sd.items())
        sig_l = tuple((k, a.data_ptr(), a._version, b.data_ptr(), b._version, s) for k, (a, b, s) in sorted(lora.items()))
        old_w, old_l = self._text_sig if self._text_sig is not None else (None, None)
        with torch.no_grad():
            if force or sig_w != old_w:
                eng.load_text_state_dict(sd)
            if force or sig_l != old_l:
                for i in range(self._text_arch.layers):
                    for which in (L.LORA_IN_PROJ, L.LORA_OUT_PROJ, L.LORA_C_FC, L.LORA_C_PROJ):
                        if (i, which) in lora:
                            a, b, s = lora[(i, which)]
                            eng.set_lora(i, which, a, b, s)
                        else:
                            eng.set_lora(i, which, None, None)
        self._text_sig = (sig_w, sig_l)
        return eng

    def _adopt_stray_lora(self) -> None:
        """/root/reference/main.py:26-27 creates its LoRA parameters on the CPU and never moves them: on a CUDA
        model `x @ self.lora_A` inside the text tower then fails on the device.  Any LoRA parameter of the text tower that lives
        on another device than the tower is moved over in place (what `model.to(device)` after the wrap would have done); the
        vision tower needs nothing - the engine uploads its operands itself."""
        dev = self.token_embedding.weight.device
        for mod in self.transformer.modules():
            lo = getattr(mod, "lora", None)
            if lo is None or not _is_lora_wrapped(mod):
                continue
            for prm in lo.parameters():
                if prm.device != dev:
                    prm.data = prm.data.to(dev)

    def encode_text(self, text: torch.Tensor) -> torch.Tensor:
        self._adopt_stray_lora()
        if self.text_on_engine:
            eng = self.sync_text_engine()
            with torch.no_grad():
                x = self.token_embedding(text).float() + self.positional_embedding.float()   # the caller's gather + add
                return eng.encode_sequence(x, text.argmax(dim=-1)).to(self.dtype)
        x = self.token_embedding(text).to(self.dtype) + self.positional_embedding.to(self.dtype)
        x = self.transformer(x.permute(1, 0, 2)).permute(1, 0, 2)
        x = self.ln_final(x)
        return x[torch.arange(x.shape[0], device=x.device), text.argmax(dim=-1)] @ self.text_projection

    def forward(self, image, text):
        i = F.normalize(self.encode_image(image).float(), dim=1)
        t = F.normalize(self.encode_text(text).float(), dim=1)
        logits = self.logit_scale.exp() * i @ t.t()
        return logits, logits.t()


# ----------------------------------------------------------------------------------------------------------------
# preprocess: callable with clip._transform's contract, computed by the CUDA kernel
# ----------------------------------------------------------------------------------------------------------------
class Preprocess:
    """preprocess(PIL.Image) -> float32 Tensor[3, R, R] on the model's device (reference call sites main.py:201,
    438, 489).  `batch(list_of_PIL)` does a whole list in one kernel launch and can emit the patch matrix directly."""

    def __init__(self, visual: VisionTransformer):
        self._visual = visual
        self.n_px = visual.input_resolution

    @staticmethod
    def _to_u8(img, device) -> torch.Tensor:
        import numpy as np
        if hasattr(img, "tensor") and isinstance(img.tensor, torch.Tensor):   # analyzer.DeviceImage: decoded on the GPU already
            t = img.tensor
            if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
                raise ValueError(f"expected a uint8 HWC RGB tensor, got {t.dtype} {tuple(t.shape)}")
            return t.to(device).contiguous()
        if hasattr(img, "convert"):
            img = img.convert("RGB")
        arr = np.array(img, dtype=np.uint8, order="C")  # owning, writable copy of the decoded pixels
        if arr.ndim != 3 or arr.shape[2] != 3:
            raise ValueError(f"expected an RGB image, got array of shape {arr.shape}")
        return torch.from_numpy(arr).to(device, non_blocking=False)

    def batch(self, images, layout: int = L.OUT_CHW_F32) -> torch.Tensor:
        eng = self._visual.engine()
        return eng.preprocess([self._to_u8(im, eng.device) for im in images], layout=layout).clone()

    def __call__(self, img) -> torch.Tensor:
        return self.batch([img])[0]


# ----------------------------------------------------------------------------------------------------------------
# tokenizer
# ----------------------------------------------------------------------------------------------------------------
def _find_bpe() -> Optional[str]:
    cands = [os.environ.get("CLIP_BPE_PATH"), os.path.expanduser("~/.cache/clip/bpe_simple_vocab_16e6.txt.gz")]
    try:
        import clip as _real  # the genuine package, if the deployment has it
        cands.append(os.path.join(os.path.dirname(_real.__file__), "bpe_simple_vocab_16e6.txt.gz"))
    except Exception:
        pass
    for c in cands:
        if c and os.path.exists(c):
            return c
    return None


def tokenize(texts: Union[str, List[str]], context_length: int = 77, truncate: bool = False,
             allow_standin: Optional[bool] = None) -> torch.Tensor:
    """clip.tokenize contract: LongTensor[N, 77] = [SOT] ids... [EOT] zero-padded; raises when too long.
    With the genuine `clip` package importable its BPE tokenizer is used.  Without it (this image: no BPE vocabulary, no
    network) the call RAISES, like the reference's would - unless the caller opts in (allow_standin=True or
    IIC_ALLOW_STANDIN_TOKENIZER=1: tests, benches) to a deterministic byte-level stand-in that keeps the same framing so
    `argmax` finds EOT.  The stand-in is only meaningful with seeded weights: with real weights it yields garbage features."""
    real = None
    try:
        import clip as _real
        # the genuine package ships its BPE vocabulary next to its sources; anything else named `clip` in sys.modules
        # (this module installed as a drop-in, a test stub) is not a tokenizer to delegate to
        if getattr(_real, "__file__", None) and _find_bpe() and getattr(_real, "tokenize", None) is not tokenize:
            real = _real
    except ImportError:
        real = None
    if real is not None:
        return real.tokenize(texts, context_length=context_length, truncate=truncate)
    if not _opted_in(allow_standin, "IIC_ALLOW_STANDIN_TOKENIZER"):
        raise RuntimeError(
            "clip.tokenize: the OpenAI `clip` package (BPE vocabulary) is not installed. Install it, or opt in to the "
            "byte-level stand-in tokenizer with tokenize(..., allow_standin=True) / IIC_ALLOW_STANDIN_TOKENIZER=1 "
            "(seeded-weight tests and benches only: with real weights it produces meaningless text features).")
    if isinstance(texts, str):
        texts = [texts]
    out = torch.zeros(len(texts), context_length, dtype=torch.long)
    for i, t in enumerate(texts):
        ids = [_SOT] + [b + 1 for b in " ".join(t.lower().split()).encode("utf-8")] + [_EOT]
        if len(ids) > context_length:
            if not truncate:
                raise RuntimeError(f"Input {t} is too long for context length {context_length}")
            ids = ids[:context_length]
            ids[-1] = _EOT
        out[i, :len(ids)] = torch.tensor(ids)
    return out


# ----------------------------------------------------------------------------------------------------------------
# load
# ----------------------------------------------------------------------------------------------------------------
def build_model(name: str = "ViT-B/16", state_dict: Optional[Dict[str, torch.Tensor]] = None, seed: int = 0) -> CLIP:
    if name not in _MODELS:
        raise RuntimeError(f"Model {name} not found; available models = {available_models()}")
    cfg = {k: v for k, v in _MODELS[name].items() if k != "file"}
    rng = torch.random.get_rng_state()
    try:
        torch.manual_seed(seed)
        model = CLIP(**cfg)
    finally:
        torch.random.set_rng_state(rng)
    if state_dict is not None:
        sd = {k: v.float() for k, v in state_dict.items()
              if k not in ("input_resolution", "context_length", "vocab_size")}
        model.load_state_dict(sd, strict=True)
    return model.eval()


def build_visual(name: str = "ViT-B/16", seed: int = 0) -> VisionTransformer:
    """Vision tower only (what the image hot path needs), seeded random weights with the upstream init scales:
    attn_std = width^-0.5, proj_std = width^-0.5 * (2*layers)^-0.5, fc_std = (2*width)^-0.5."""
    cfg = _MODELS[name]
    rng = torch.random.get_rng_state()
    try:
        torch.manual_seed(seed)
        w, layers = cfg["vision_width"], cfg["vision_layers"]
        v = VisionTransformer(cfg["image_resolution"], cfg["vision_patch_size"], w, layers, w // 64, cfg["embed_dim"])
        attn_std, fc_std = w ** -0.5, (2 * w) ** -0.5
        proj_std = attn_std * (2 * layers) ** -0.5
        for b in v.transformer.resblocks:
            nn.init.normal_(b.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(b.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(b.mlp.c_fc.weight, std=fc_std)
            nn.init.normal_(b.mlp.c_proj.weight, std=proj_std)
            for t in (b.attn.in_proj_bias, b.attn.out_proj.bias, b.mlp.c_fc.bias, b.mlp.c_proj.bias):
                nn.init.normal_(t, std=0.02)
    finally:
        torch.random.set_rng_state(rng)
    return v.eval()


def _checkpoint_state_dict(path: str) -> Dict[str, torch.Tensor]:
    """OpenAI checkpoint (TorchScript archive) or a plain state dict.  Parse errors PROPAGATE: a corrupt or partly
    written file must not silently turn into random weights."""
    try:
        return torch.jit.load(path, map_location="cpu").state_dict()
    except RuntimeError as jit_err:
        try:
            sd = torch.load(path, map_location="cpu")
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(f"cannot read CLIP checkpoint {path}: not a TorchScript archive ({jit_err}) and torch.load "
                               f"failed ({e})") from e
    if isinstance(sd, dict):
        sd = sd.get("state_dict", sd)
    if not isinstance(sd, dict) or not sd:
        raise RuntimeError(f"CLIP checkpoint {path} does not contain a state dict")
    return sd


def load(name: str = "ViT-B/16", device: Union[str, torch.device, None] = None, jit: bool = False,
         download_root: Optional[str] = None, state_dict: Optional[Dict[str, torch.Tensor]] = None,
         operand_dtype: Optional[str] = None, allow_random_init: Optional[bool] = None):
    """clip.load contract: returns (model.eval(), preprocess).

    Weights: `state_dict` if given; else the OpenAI checkpoint at `download_root or ~/.cache/clip/<file>` if it is
    on disk (nothing is ever downloaded; a file that is there but cannot be parsed RAISES).  With neither, the call
    raises like the reference's `clip.load` would when its download fails - unless the caller opts in
    (allow_random_init=True or IIC_ALLOW_RANDOM_INIT=1: tests, benches) to a seeded random initialisation with the
    upstream init scales, which is announced with a warning.  A misconfigured deployment therefore never serves
    classifications from random weights silently.
    Parameters stay fp32 (master copy); the engine converts matmul weights to the 16-bit operand dtype on upload
    (`operand_dtype`: "f16" default | "bf16").  `jit` is ignored."""
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    if name not in _MODELS:
        raise RuntimeError(f"Model {name} not found; available models = {available_models()}")
    if state_dict is None:
        path = os.path.join(download_root or os.path.expanduser("~/.cache/clip"), _MODELS[name]["file"])
        if os.path.isfile(path):
            state_dict = _checkpoint_state_dict(path)
        elif _opted_in(allow_random_init, "IIC_ALLOW_RANDOM_INIT"):
            warnings.warn(f"clip.load({name!r}): no checkpoint at {path}; using SEEDED RANDOM weights "
                          f"(allow_random_init) - outputs are meaningless outside tests and benches", stacklevel=2)
        else:
            raise RuntimeError(
                f"clip.load({name!r}): checkpoint {path} not found and nothing is downloaded (no network). Put the OpenAI "
                f"checkpoint there, pass state_dict=..., or opt in to seeded random weights with allow_random_init=True / "
                f"IIC_ALLOW_RANDOM_INIT=1 (tests and benches only).")
    model = build_model(name, state_dict).to(device)
    if operand_dtype is not None:
        model.visual.operand_dtype = operand_dtype
    for p in model.parameters():
        p.requires_grad_(True)
    return model, Preprocess(model.visual)
