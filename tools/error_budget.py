"""Where does the bf16 error come from?  Emulates bf16 rounding at chosen points of the fp32 oracle forward (CPU) and
reports the logit error each one causes.  Dev tool (uses oracle/ + tests/golden)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import torch.nn.functional as F
from _common import golden_npz, oracle_model

def r(t, on):
    return t.to(torch.bfloat16).to(torch.float32) if on else t

def forward(m, x, pts):
    v = m.visual
    x = r(x, "input" in pts)
    x = F.conv2d(x, v.conv1.weight, stride=v.conv1.stride)
    x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
    x = torch.cat([v.class_embedding.expand(x.shape[0], 1, -1), x], 1) + v.positional_embedding
    x = F.layer_norm(x, (x.shape[-1],), v.ln_pre.weight, v.ln_pre.bias, 1e-5)
    B, T, d = x.shape
    H = d // 64
    for blk in v.transformer.resblocks:
        y = r(F.layer_norm(x, (d,), blk.ln_1.weight, blk.ln_1.bias, 1e-5), "ln" in pts)
        qkv = r(y @ blk.attn.in_proj_weight.t() + blk.attn.in_proj_bias, "qkv" in pts)
        q, k, vv = qkv.view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
        p = (q @ k.transpose(-1, -2) / 8.0).softmax(-1)
        if "probs" in pts:
            mx = p.max(-1, keepdim=True).values   # kernel rounds exp(s - max) (un-normalised) to bf16
            p = r(p / mx, True) * mx
        o = r((p @ vv).permute(0, 2, 1, 3).reshape(B, T, d), "attn_out" in pts)
        x = x + o @ blk.attn.out_proj.weight.t() + blk.attn.out_proj.bias
        y = r(F.layer_norm(x, (d,), blk.ln_2.weight, blk.ln_2.bias, 1e-5), "ln" in pts)
        h = y @ blk.mlp.c_fc.weight.t() + blk.mlp.c_fc.bias
        h = r(h * torch.sigmoid(1.702 * h), "gelu" in pts)
        x = x + h @ blk.mlp.c_proj.weight.t() + blk.mlp.c_proj.bias
        if "resid" in pts: x = r(x, True)
    x = F.layer_norm(x[:, 0], (d,), v.ln_post.weight, v.ln_post.bias, 1e-5)
    return x @ v.proj

def main():
    torch.set_num_threads(8)
    m = oracle_model()
    n = int(os.environ.get("N", "24"))
    crops = torch.from_numpy(golden_npz("crops_u8.npz")["crops"][:n])
    mean = torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(1, 3, 1, 1)
    std = torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(1, 3, 1, 1)
    x = (crops.permute(0, 3, 1, 2).float() / 255 - mean) / std
    text = torch.from_numpy(golden_npz("text_features.npz")["text"])
    ref = torch.from_numpy(golden_npz("ref_shipped.npz")["logits"][:n])
    def logits(e):
        f = e / e.norm(dim=-1, keepdim=True)
        return 100 * f @ text.t()
    with torch.no_grad():
        base = logits(forward(m, x, set()))
        print("fp32 restatement vs golden: max", (base - ref).abs().max().item())
        for pts in (["input"], ["ln"], ["qkv"], ["probs"], ["attn_out"], ["gelu"], ["resid"],
                    ["ln", "qkv", "probs", "attn_out", "gelu"], ["input", "ln", "qkv", "probs", "attn_out", "gelu"]):
            d = (logits(forward(m, x, set(pts))) - base).abs()
            print(f"{'+'.join(pts):45s} max {d.max():.4f} rms {d.pow(2).mean().sqrt():.4f}")

main()
