"""ORACLE (test infrastructure only): plain-PyTorch restatement of the parts of /root/reference/main.py that sit on
the hot path, for use where /root/reference is not mounted (the GPU box).  Every function cites the lines it follows.
oracle/gen_golden.py asserts, in the build container, that these restatements reproduce the reference's own outputs
on all 151 images before any fixture is written.

Pinned by: tests/golden/ref_*.npz (outputs of the unmodified reference code) and the loader known-answer
"48 loaded / 96 missing" (SURVEY F7).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn as nn

GROUP_ORDER = ("styles", "characteristics", "materials", "colors", "room_types")  # main.py:289-295 dict order
N_INTERIOR = 11  # main.py:185


# ------------------------------------------------------------------------------------------------- LoRA (main.py:19-74)
class LoRALayer(nn.Module):
    """main.py:19-31"""

    def __init__(self, in_dim, out_dim, rank=4, alpha=8):
        super().__init__()
        self.rank, self.alpha = rank, alpha
        self.lora_A = nn.Parameter(torch.randn(in_dim, rank) * 0.02)
        self.lora_B = nn.Parameter(torch.zeros(rank, out_dim))
        self.scaling = self.alpha / self.rank

    def forward(self, x):
        return (x @ self.lora_A @ self.lora_B) * self.scaling


class LoRALinear(nn.Module):
    """main.py:34-58: linear(x) + lora(x); .weight/.bias proxies keep nn.MultiheadAttention working, which is also
    why an out_proj LoRA never contributes (F4)."""

    def __init__(self, linear_module: nn.Linear, rank=4, alpha=8):
        super().__init__()
        self.linear = linear_module
        self.lora = LoRALayer(linear_module.in_features, linear_module.out_features, rank=rank, alpha=alpha)

    def forward(self, x):
        return self.linear(x) + self.lora(x)

    weight = property(lambda self: self.linear.weight)
    bias = property(lambda self: self.linear.bias)
    in_features = property(lambda self: self.linear.in_features)
    out_features = property(lambda self: self.linear.out_features)


def replace_linears_with_lora(module: nn.Module, rank=4, alpha=8, names=None, parent="") -> List[str]:
    """main.py:62-74: recursive in-place wrap of every nn.Linear child."""
    names = [] if names is None else names
    for name, child in list(module.named_children()):
        full = f"{parent}.{name}" if parent else name
        if isinstance(child, nn.Linear):
            setattr(module, name, LoRALinear(child, rank=rank, alpha=alpha))
            names.append(full)
        else:
            replace_linears_with_lora(child, rank, alpha, names, full)
    return names


def load_lora_state(model: nn.Module, ckpt: Dict[str, torch.Tensor]) -> Tuple[int, List[str]]:
    """main.py:86-113 on an already-loaded dict: exact name, else first key with k.endswith(name) or
    name.endswith(k); non-strict."""
    keys, loaded, missing = list(ckpt.keys()), 0, []
    for name, param in model.named_parameters():
        if "lora" not in name:
            continue
        if name in ckpt:
            param.data = ckpt[name].to(param.device)
            loaded += 1
            continue
        hit = next((k for k in keys if k.endswith(name) or name.endswith(k)), None)
        if hit is not None:
            param.data = ckpt[hit].to(param.device)
            loaded += 1
        else:
            missing.append(name)
    return loaded, missing


def loader_kat(ref_main, ckpt_path: str) -> Dict[str, int]:
    """Known answers derivable from the repo's artefacts: 72 wrapped layers; the shipped checkpoint loads
    48 tensors and leaves 96 missing; every visual lora_B stays zero."""
    from oracle import clip_ref
    model, _ = clip_ref.load("ViT-B/16")
    wrapped = ref_main.replace_linears_with_lora(model, rank=4, alpha=8)
    loaded, missing = ref_main.load_lora_weights_to_model(model, ckpt_path, strict_match=False)
    visual_b_zero = all(bool((p == 0).all()) for n, p in model.named_parameters()
                        if n.startswith("visual.") and n.endswith("lora_B"))
    return {"wrapped": len(wrapped), "loaded": loaded, "missing": len(missing), "visual_lora_B_all_zero": int(visual_b_zero)}


def seed_vision_lora(model: nn.Module, seed: int = 1234, b_std: float = 0.02) -> None:
    """Make the vision-tower LoRA non-zero, deterministically: lora_A ~ N(0, 0.02^2) (main.py:26), lora_B ~
    N(0, b_std^2) rounded to bf16-representable values (the engine holds lora_B^T in bf16), for c_fc, c_proj AND
    out_proj of every visual block - the latter must stay without effect (F4)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if not n.startswith("visual.") or "lora" not in n:
                continue
            if n.endswith("lora_A"):
                p.data = torch.randn(p.shape, generator=g) * 0.02
            elif n.endswith("lora_B"):
                p.data = (torch.randn(p.shape, generator=g) * b_std).to(torch.bfloat16).to(torch.float32)


def weights_checksum(model: nn.Module) -> Dict[str, float]:
    sd = model.state_dict()
    keys = ["visual.conv1.weight", "visual.proj", "visual.transformer.resblocks.11.mlp.c_fc.weight",
            "visual.transformer.resblocks.0.attn.in_proj_weight", "text_projection", "token_embedding.weight"]
    return {k: float(sd[k].double().abs().sum()) for k in keys if k in sd}


# ------------------------------------------------------------------------------------------------- heads
def detector_decision(logits40: torch.Tensor, categories: Sequence[str], threshold: float = 0.3):
    """main.py:208-222 given logits = 100 * f . T_det^T for one image:
    softmax over the 40 prompts, top-1, sum of the first 11 vs the rest, threshold on the top-1 confidence."""
    p = logits40.softmax(dim=-1)
    top_conf, top_idx = p.topk(1)
    interior = p[:N_INTERIOR].sum().item()
    non_interior = p[N_INTERIOR:].sum().item()
    is_interior = interior > non_interior and top_conf.item() > threshold
    return is_interior, interior, categories[top_idx.item()]


def group_topk(logits397: torch.Tensor, groups: Dict[str, List[str]], k: int = 5) -> Dict[str, List[Tuple[str, float]]]:
    """main.py:455-459 == 505-509 given logits = 100 * f . T_g^T for one image, groups concatenated in GROUP_ORDER:
    per group softmax, topk(min(5, |g|)), (label, prob) pairs in descending order."""
    out, off = {}, 0
    for g in GROUP_ORDER:
        n = len(groups[g])
        p = logits397[off:off + n].softmax(dim=-1)
        vals, inds = p.topk(min(k, n))
        out[g] = [(groups[g][i], v.item()) for v, i in zip(vals, inds)]
        off += n
    return out
