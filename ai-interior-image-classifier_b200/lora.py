"""LoRA adapter API with the reference's names, parameter naming and checkpoint layout.

Mirrors /root/reference/main.py:19-113 (inference: rank 4 / alpha 8) and train_lora.py:15-44 (training: rank 16 /
alpha 32 / dropout) so that code written against the reference keeps working and `lora_models/*.pth` files
round-trip byte-compatibly:

  * wrapped layer `X` exposes parameters `X.linear.weight`, `X.linear.bias`, `X.lora.lora_A` (in, r), `X.lora.lora_B`
    (r, out);  lora_A ~ N(0, 0.02^2), lora_B = 0;  forward = linear(x) + (x @ A @ B) * (alpha / rank);
  * checkpoint = plain dict {parameter name: fp32 CPU tensor} of every parameter whose name contains 'lora',
    written with torch.save; keys may carry a `clip_model.` prefix (train_lora.py writes through LoRACLIPWrapper);
  * loading is non-strict with suffix matching (exact name first, else the first checkpoint key k with
    k.endswith(name) or name.endswith(k)), returns (loaded, missing).

On the VISION tower the wrapped modules are only parameter holders: the engine reads (A, B, alpha/rank) and fuses
the update into the GEMM tile (csrc/gemm_sm100.cuh).  On the TEXT tower `LoRALinear.forward` below is what runs.
One deliberate difference from main.py:26-27: the LoRA parameters are created on the wrapped layer's device and
dtype (the reference leaves them on the CPU, which is why its LoRA path only runs there - SURVEY F8).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn


class LoRALayer(nn.Module):
    def __init__(self, in_dim: int, out_dim: int, rank: int = 4, alpha: float = 8, dropout: float = 0.0, device=None,
                 dtype=None):
        super().__init__()
        self.in_dim, self.out_dim, self.rank, self.alpha = in_dim, out_dim, rank, alpha
        self.lora_A = nn.Parameter(torch.randn(in_dim, rank, device=device, dtype=dtype) * 0.02)
        self.lora_B = nn.Parameter(torch.zeros(rank, out_dim, device=device, dtype=dtype))
        self.scaling = self.alpha / self.rank
        self.dropout = nn.Dropout(dropout) if dropout > 0.0 else nn.Identity()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.dropout((x @ self.lora_A @ self.lora_B) * self.scaling)


class LoRALinear(nn.Module):
    def __init__(self, linear_module: nn.Linear, rank: int = 4, alpha: float = 8, dropout: float = 0.0):
        super().__init__()
        self.linear = linear_module
        w = linear_module.weight
        self.lora = LoRALayer(linear_module.in_features, linear_module.out_features, rank=rank, alpha=alpha,
                              dropout=dropout, device=w.device, dtype=w.dtype)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.linear(x) + self.lora(x)

    # nn.MultiheadAttention reads out_proj.weight / out_proj.bias directly: keep them reachable (main.py:45-51)
    @property
    def weight(self):
        return self.linear.weight

    @property
    def bias(self):
        return self.linear.bias

    @property
    def in_features(self):
        return self.linear.in_features

    @property
    def out_features(self):
        return self.linear.out_features


def replace_linears_with_lora(module: nn.Module, rank: int = 4, alpha: float = 8, replaced_names: Optional[List[str]] = None,
                              parent_name: str = "") -> List[str]:
    """Recursively wrap every nn.Linear child in place; returns the dotted names wrapped (main.py:62-74).
    On a CLIP model: attn.out_proj, mlp.c_fc, mlp.c_proj of every block of both towers = 72 layers; the packed QKV
    projection is a bare Parameter of nn.MultiheadAttention and is never wrapped."""
    if replaced_names is None:
        replaced_names = []
    for name, child in list(module.named_children()):
        full = f"{parent_name}.{name}" if parent_name else name
        if isinstance(child, nn.Linear):
            setattr(module, name, LoRALinear(child, rank=rank, alpha=alpha))
            replaced_names.append(full)
        elif not isinstance(child, LoRALinear):
            replace_linears_with_lora(child, rank, alpha, replaced_names, full)
    return replaced_names


def lora_state_dict(model: nn.Module) -> Dict[str, torch.Tensor]:
    return {n: p.detach().to("cpu", torch.float32).contiguous() for n, p in model.named_parameters() if "lora" in n}


def save_lora_weights(model: nn.Module, path: str) -> None:
    sd = lora_state_dict(model)
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    torch.save(sd, path)
    print(f"Zapisano {len(sd)} parametrów LoRA do {path}")


def load_lora_weights_to_model(model: nn.Module, path: str, strict_match: bool = False) -> Tuple[int, List[str]]:
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    ckpt = torch.load(path, map_location="cpu")
    keys = list(ckpt.keys())
    loaded, missing = 0, []
    for name, param in model.named_parameters():
        if "lora" not in name:
            continue
        src = ckpt.get(name)
        if src is None:
            hit = next((k for k in keys if k.endswith(name) or name.endswith(k)), None)
            src = ckpt[hit] if hit is not None else None
        if src is None:
            missing.append(name)
            continue
        param.data = src.to(param.device)
        loaded += 1
    print(f"Wczytano {loaded} LoRA parametrów z {path}. Brakujących: {len(missing)}")
    if strict_match and missing:
        raise RuntimeError(f"Nie wczytano wszystkich LoRA parametrów, brak: {missing[:10]}")
    return loaded, missing
