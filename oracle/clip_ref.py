"""ORACLE (test infrastructure, never the product path): CPU/PyTorch fp32 restatement of OpenAI CLIP as the
reference calls it.

Only `tests/`, `__graft_entry__.smoke()`, `bench.py`'s cpu_baseline / `--impl reference` leg and `oracle/` scripts
may import this module.  The product package never does.

Why a restatement: the reference's arithmetic for this path lives in the third-party `clip` package
(`openai-clip` / `clip-anytorch`, both listed UNPINNED in /root/reference/python-worker/requirements.txt:3,17 and
not vendored under /root/reference; neither the package, its BPE vocabulary nor any pretrained weights exist in
this image and there is no network).  This file restates the published architecture of `clip/model.py`
(`CLIP`, `VisionTransformer`, `ResidualAttentionBlock`, `QuickGELU`, fp32 `LayerNorm`), `clip.load`'s return
contract and `clip._transform`, with the upstream state-dict names, so that /root/reference/main.py and
train_lora.py import and run UNMODIFIED on top of it (`install_clip_stub()`), which makes the LoRA wrap, the
checkpoint loader, both heads and the training step genuinely the reference's own code.

Parity pin: the reference holds no golden vectors for this path ("parity unpinned" by the
reference's own tests).  The restatement is pinned instead by (i) an independent implementation of the same
published architecture, `transformers.CLIPVisionModelWithProjection` / `CLIPTextModelWithProjection`, fed the
same tensors (tests/test_oracle.py, fp32 agreement <= 1e-4), and (ii) outputs of the reference's own code run in
the build container on top of it, committed under tests/golden/ by oracle/gen_golden.py.

Call sites in the reference this serves: main.py:152,180-182,204,241,307-309,444,503; train_lora.py:174,233,237,241.
"""
from __future__ import annotations

import math
import sys
import types
from collections import OrderedDict
from typing import Dict, List, Tuple, Union

import torch
import torch.nn as nn

# (image_size, patch, vision width, vision layers, embed_dim, text width, text heads, text layers)
ARCHS = {
    "ViT-B/16": dict(image_resolution=224, vision_patch_size=16, vision_width=768, vision_layers=12, embed_dim=512,
                     context_length=77, vocab_size=49408, transformer_width=512, transformer_heads=8,
                     transformer_layers=12),
    "ViT-L/14@336px": dict(image_resolution=336, vision_patch_size=14, vision_width=1024, vision_layers=24,
                           embed_dim=768, context_length=77, vocab_size=49408, transformer_width=768,
                           transformer_heads=12, transformer_layers=12),
}
SOT_TOKEN, EOT_TOKEN = 49406, 49407


class LayerNorm(nn.LayerNorm):
    """upstream: computes in fp32 whatever the input dtype, casts back."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return super().forward(x.type(torch.float32)).type(x.dtype)


class QuickGELU(nn.Module):
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model: int, n_head: int, attn_mask: torch.Tensor = None):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, d_model * 4)),
            ("gelu", QuickGELU()),
            ("c_proj", nn.Linear(d_model * 4, d_model)),
        ]))
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask

    def attention(self, x: torch.Tensor) -> torch.Tensor:
        mask = self.attn_mask.to(dtype=x.dtype, device=x.device) if self.attn_mask is not None else None
        return self.attn(x, x, x, need_weights=False, attn_mask=mask)[0]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = x + self.attention(self.ln_1(x))
        x = x + self.mlp(self.ln_2(x))
        return x


class Transformer(nn.Module):
    def __init__(self, width: int, layers: int, heads: int, attn_mask: torch.Tensor = None):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask) for _ in range(layers)])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.resblocks(x)


class VisionTransformer(nn.Module):
    def __init__(self, input_resolution: int, patch_size: int, width: int, layers: int, heads: int, output_dim: int):
        super().__init__()
        self.input_resolution, self.output_dim = input_resolution, output_dim
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.conv1(x)                                   # [B, width, g, g]
        x = x.reshape(x.shape[0], x.shape[1], -1)           # [B, width, g*g]
        x = x.permute(0, 2, 1)                              # [B, g*g, width]
        cls = self.class_embedding.to(x.dtype) + torch.zeros(x.shape[0], 1, x.shape[-1], dtype=x.dtype, device=x.device)
        x = torch.cat([cls, x], dim=1)                      # [B, T, width]
        x = x + self.positional_embedding.to(x.dtype)
        x = self.ln_pre(x)
        x = x.permute(1, 0, 2)                              # [T, B, width]
        x = self.transformer(x)
        x = x.permute(1, 0, 2)
        x = self.ln_post(x[:, 0, :])
        if self.proj is not None:
            x = x @ self.proj
        return x


class CLIP(nn.Module):
    def __init__(self, embed_dim: int, image_resolution: int, vision_layers: int, vision_width: int,
                 vision_patch_size: int, context_length: int, vocab_size: int, transformer_width: int,
                 transformer_heads: int, transformer_layers: int):
        super().__init__()
        self.context_length = context_length
        self.visual = VisionTransformer(image_resolution, vision_patch_size, vision_width, vision_layers,
                                        vision_width // 64, embed_dim)
        self.transformer = Transformer(transformer_width, transformer_layers, transformer_heads,
                                       attn_mask=self.build_attention_mask())
        self.vocab_size = vocab_size
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(context_length, transformer_width))
        self.ln_final = LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))
        self.initialize_parameters()

    def initialize_parameters(self):
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        for tower in (self.transformer, self.visual.transformer):
            proj_std = (tower.width ** -0.5) * ((2 * tower.layers) ** -0.5)
            attn_std = tower.width ** -0.5
            fc_std = (2 * tower.width) ** -0.5
            for block in tower.resblocks:
                nn.init.normal_(block.attn.in_proj_weight, std=attn_std)
                nn.init.normal_(block.attn.out_proj.weight, std=proj_std)
                nn.init.normal_(block.mlp.c_fc.weight, std=fc_std)
                nn.init.normal_(block.mlp.c_proj.weight, std=proj_std)
        nn.init.normal_(self.text_projection, std=self.transformer.width ** -0.5)

    def build_attention_mask(self) -> torch.Tensor:
        mask = torch.empty(self.context_length, self.context_length)
        mask.fill_(float("-inf"))
        mask.triu_(1)
        return mask

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    def encode_image(self, image: torch.Tensor) -> torch.Tensor:
        return self.visual(image.type(self.dtype))

    def encode_text(self, text: torch.Tensor) -> torch.Tensor:
        x = self.token_embedding(text).type(self.dtype)
        x = x + self.positional_embedding.type(self.dtype)
        x = x.permute(1, 0, 2)
        x = self.transformer(x)
        x = x.permute(1, 0, 2)
        x = self.ln_final(x).type(self.dtype)
        # features of the end-of-text token (the highest id in each sequence)
        x = x[torch.arange(x.shape[0]), text.argmax(dim=-1)] @ self.text_projection
        return x

    def forward(self, image, text):
        image_features = self.encode_image(image)
        text_features = self.encode_text(text)
        image_features = image_features / image_features.norm(dim=1, keepdim=True)
        text_features = text_features / text_features.norm(dim=1, keepdim=True)
        logit_scale = self.logit_scale.exp()
        logits_per_image = logit_scale * image_features @ text_features.t()
        return logits_per_image, logits_per_image.t()


# ----------------------------------------------------------------------------------------------------------------
# seeded weights (no pretrained checkpoint is reachable offline)
# ----------------------------------------------------------------------------------------------------------------
def round_to_bf16_(t: torch.Tensor) -> torch.Tensor:
    t.data = t.data.to(torch.bfloat16).to(torch.float32)
    return t


def build_model(name: str = "ViT-B/16", seed: int = 0, bf16_representable: bool = True,
                released_logit_scale: bool = True) -> CLIP:
    """Deterministic random-init CLIP with the upstream init scales.

    bf16_representable: round every matmul weight of the VISION tower (conv1, in_proj, out_proj, c_fc, c_proj) to
    a bf16-representable fp32 value, so the fp32 oracle and the bf16 engine hold numerically identical weights and
    parity measures arithmetic, not weight quantisation (BASELINE config 3: "weights ... cast to bf16").
    LayerNorm affine parameters and biases get small seeded perturbations so that no term of the forward is
    trivially 1 or 0 (a zero bias would hide a dropped bias add)."""
    gen_state = torch.random.get_rng_state()
    try:
        torch.manual_seed(seed)
        model = CLIP(**ARCHS[name])
        g = torch.Generator().manual_seed(seed + 1)
        with torch.no_grad():
            for n, p in model.named_parameters():
                if n.endswith(("ln_1.weight", "ln_2.weight", "ln_pre.weight", "ln_post.weight", "ln_final.weight")):
                    p.add_(0.05 * torch.randn(p.shape, generator=g))
                elif n.endswith(".bias") or n.endswith("in_proj_bias"):
                    p.copy_(0.02 * torch.randn(p.shape, generator=g))
            if released_logit_scale:
                model.logit_scale.fill_(math.log(100.0))
            if bf16_representable:
                v = model.visual
                round_to_bf16_(v.conv1.weight)
                for blk in v.transformer.resblocks:
                    for p in (blk.attn.in_proj_weight, blk.attn.out_proj.weight, blk.mlp.c_fc.weight,
                              blk.mlp.c_proj.weight):
                        round_to_bf16_(p)
    finally:
        torch.random.set_rng_state(gen_state)
    return model.eval()


# ----------------------------------------------------------------------------------------------------------------
# clip.load / clip.tokenize / clip._transform stand-ins
# ----------------------------------------------------------------------------------------------------------------
def _convert_image_to_rgb(image):
    return image.convert("RGB")


def transform(n_px: int):
    """clip._transform: the real torchvision + Pillow pipeline (these ARE importable here)."""
    from torchvision.transforms import CenterCrop, Compose, InterpolationMode, Normalize, Resize, ToTensor
    return Compose([
        Resize(n_px, interpolation=InterpolationMode.BICUBIC),
        CenterCrop(n_px),
        _convert_image_to_rgb,
        ToTensor(),
        Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)),
    ])


def tokenize(texts: Union[str, List[str]], context_length: int = 77, truncate: bool = False) -> torch.Tensor:
    """Deterministic stand-in for clip.tokenize (the BPE vocabulary is not available offline): lower-cased UTF-8
    bytes -> ids in [1, 256], wrapped in the real SOT/EOT ids and zero padded, so that `text.argmax(-1)` picks the
    EOT position exactly as upstream.  The text tower only manufactures the fixed [L, E] label matrix the image
    head scores against; any injective, deterministic tokenisation serves parity."""
    if isinstance(texts, str):
        texts = [texts]
    out = torch.zeros(len(texts), context_length, dtype=torch.long)
    for i, t in enumerate(texts):
        ids = [SOT_TOKEN] + [b + 1 for b in " ".join(t.lower().split()).encode("utf-8")] + [EOT_TOKEN]
        if len(ids) > context_length:
            if not truncate:
                raise RuntimeError(f"Input {t} is too long for context length {context_length}")
            ids = ids[:context_length]
            ids[-1] = EOT_TOKEN
        out[i, :len(ids)] = torch.tensor(ids)
    return out


_MODEL_SEEDS: Dict[str, int] = {}
_load_calls = 0


def load(name: str, device: Union[str, torch.device] = "cpu", jit: bool = False, download_root: str = None,
         seed: int = None) -> Tuple[CLIP, object]:
    """clip.load contract: (model.eval(), preprocess); fp32 on CPU.  Every call with the same name returns the SAME
    seeded weights (like re-loading one checkpoint file) - the reference loads "ViT-B/16" twice (main.py:152, 241)."""
    if name not in ARCHS:
        raise RuntimeError(f"Model {name} not found; available models = {list(ARCHS)}")
    global _load_calls
    _load_calls += 1
    model = build_model(name, seed=_MODEL_SEEDS.get(name, 0) if seed is None else seed)
    model = model.to(device)
    if str(device) != "cpu":
        # upstream converts weights to fp16 on CUDA; the oracle is only ever run on CPU
        raise RuntimeError("oracle clip stub is CPU/fp32 only (the reference's LoRA path only works there, SURVEY F8)")
    return model, transform(model.visual.input_resolution)


def available_models() -> List[str]:
    return list(ARCHS)


def install_clip_stub() -> types.ModuleType:
    """Put a `clip` module exposing load / tokenize / available_models into sys.modules so the reference's
    `import clip` resolves to this restatement."""
    mod = types.ModuleType("clip")
    mod.load, mod.tokenize, mod.available_models = load, tokenize, available_models
    mod._transform = transform
    mod.__doc__ = "oracle stub of OpenAI CLIP (see oracle/clip_ref.py)"
    sys.modules["clip"] = mod
    return mod
