"""CLIP's text tower on the CUDA engine (model.encode_text, /root/reference/main.py:181, 308;
train_lora.py:237; main_API.py:161) - the same GEMM / LayerNorm / tcgen05 attention kernels with T = 77, width 512, a
causal mask and a LIVE text-side LoRA on mlp.c_fc / mlp.c_proj (the placement of the shipped lora_models/*.pth
checkpoints).  Checked against the fp32 CPU oracle wrapped by the restated reference LoRA code."""
import copy

import pytest
import torch

from _common import golden_json, oracle_model, oracle_state_dict

pytestmark = pytest.mark.gpu


def _bf16(t):
    return t.to(torch.bfloat16)


@pytest.mark.parametrize("T,B,H", [(77, 5, 8), (77, 64, 12), (200, 3, 8), (333, 2, 4), (16, 2, 8)])
def test_causal_attention(iic, T, B, H):
    """causal mask inside the tcgen05 attention kernel (single key block for T = 77; several blocks, some of them entirely
    in a query's future, for longer sequences)"""
    arch = iic.VisionArch(image_size=224, patch_size=16, width=512, layers=1, heads=8, embed_dim=512, seq_tokens=77, causal=True)
    eng = iic.Engine(arch, "cuda:0", operand_dtype="bf16")
    d = H * 64
    g = torch.Generator(device="cuda").manual_seed(21)
    qkv = _bf16(torch.randn(B * T, 3 * d, device="cuda", generator=g))
    out = eng.op_attention(qkv, B, T, H, impl=2)
    q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True).permute(0, 2, 1, 3).reshape(B * T, d)
    assert torch.allclose(out.float(), ref, rtol=2 ** -6, atol=2e-2), (out.float() - ref).abs().max()
    assert ((out.float() - ref).norm() / ref.norm()).item() < 8e-3


@pytest.mark.parametrize("mode", ["f16", "bf16"])
def test_encode_text_on_engine_matches_oracle(iic, mode):
    from oracle import clip_ref, ref_semantics as RS
    lab = golden_json("labels.json")
    prompts = list(lab["detector"][:12])
    for g in lab["group_order"]:
        tmpl = "{}" if g == "room_types" else "wnętrze z {}"            # main.py:302-305
        prompts += [tmpl.format(l) for l in lab["groups"][g][:10]]
    tokens = clip_ref.tokenize(prompts)
    # ---- oracle: reference LoRA wrap (main.py:62-74) of the whole model, text-side LoRA seeded non-zero ----
    om = copy.deepcopy(oracle_model())
    RS.replace_linears_with_lora(om, rank=4, alpha=8)
    gen = torch.Generator().manual_seed(77)
    with torch.no_grad():
        for n, p in om.named_parameters():
            if n.startswith("transformer.") and n.endswith("lora_A"):
                p.data = torch.randn(p.shape, generator=gen) * 0.02
            elif n.startswith("transformer.") and n.endswith("lora_B"):
                p.data = (torch.randn(p.shape, generator=gen) * 0.01).to(torch.bfloat16).float()
        ref = om.encode_text(tokens)
        base = oracle_model().encode_text(tokens)
    assert ((ref - base).norm() / base.norm()).item() > 0.02          # the text LoRA matters: cannot pass by skipping it
    # ---- product: same tensors, text tower on the engine ----
    model, _ = iic.load("ViT-B/16", device="cuda", state_dict=oracle_state_dict(), operand_dtype=mode)
    iic.replace_linears_with_lora(model, rank=4, alpha=8)
    src = {n: p for n, p in om.named_parameters() if "lora" in n}
    for n, p in model.named_parameters():
        if n in src:
            p.data = src[n].detach().clone().to(p.device)
    with torch.no_grad():
        torch_path = model.encode_text(tokens.cuda())                    # PyTorch fp32 (default)
        model.text_on_engine = True
        got = model.encode_text(tokens.cuda())
    assert got.shape == ref.shape == (len(prompts), 512)
    assert torch.allclose(torch_path.cpu(), ref, rtol=1e-3, atol=1e-3)  # the mirror module itself agrees with the oracle
    cos = torch.nn.functional.cosine_similarity(got.cpu().double(), ref.double(), dim=-1)
    rel = ((got.cpu() - ref).norm() / ref.norm()).item()
    print(f"\n[text tower {mode}] cos min {cos.min():.6f}, relative embedding error {rel:.2e}")
    assert cos.min().item() >= 0.999
    assert rel < (3e-3 if mode == "f16" else 2e-2)
    # label logits against a fixed unit image embedding move by less than the image-side logit bar
    f = torch.nn.functional.normalize(torch.randn(4, 512, generator=gen), dim=-1)
    l_ref = 100.0 * f @ torch.nn.functional.normalize(ref, dim=-1).t()
    l_got = 100.0 * f @ torch.nn.functional.normalize(got.cpu(), dim=-1).t()
    # (bf16 operands: 8-bit mantissa, 7e-3 relative embedding error -> up to ~0.15 on a 100 * cos logit; fp16 meets the 2e-2 bar)
    assert (l_ref - l_got).abs().max().item() < (2e-2 if mode == "f16" else 0.25)
