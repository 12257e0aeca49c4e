"""BASELINE configs[4]: the larger encoder, ViT-L/14 at 336x336 (T = 577, d = 1024, 24 layers, patch K = 588 -> 592),
LoRA rank 16 / alpha 32 (train_lora.py:168 defaults) on the vision MLPs.  Same kernels, different shapes: parity of the
embedding against the oracle's fp32 CPU forward on a small batch (the oracle needs ~5 s per image here)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["f16", "bf16"])
def test_vit_l14_336_embedding_parity(iic, mode):
    from oracle import clip_ref, ref_semantics as RS
    om = clip_ref.build_model("ViT-L/14@336px", seed=3)
    sd = {k: v.clone() for k, v in om.state_dict().items()}
    RS.replace_linears_with_lora(om.visual, rank=16, alpha=32)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for n, p in om.visual.named_parameters():
            if n.endswith("lora_A"):
                p.data = torch.randn(p.shape, generator=g) * 0.02
            elif n.endswith("lora_B"):
                p.data = (torch.randn(p.shape, generator=g) * 0.01).to(torch.bfloat16).float()
    B = 2
    u8 = torch.randint(0, 256, (B, 336, 336, 3), dtype=torch.uint8, generator=g)
    mean = torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(1, 3, 1, 1)
    std = torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(1, 3, 1, 1)
    x = (u8.permute(0, 3, 1, 2).float() / 255 - mean) / std
    with torch.no_grad():
        ref = om.encode_image(x)
    model, _ = iic.load("ViT-L/14@336px", device="cuda", state_dict=sd, operand_dtype=mode)
    iic.replace_linears_with_lora(model.visual, rank=16, alpha=32)
    src = dict(om.visual.named_parameters())
    for n, p in model.visual.named_parameters():
        if "lora" in n:
            p.data = src[n].detach().clone().to(p.device)
    eng = model.visual.sync_engine()
    assert eng.dims.tokens == 577 and eng.dims.patch_k == 588 and eng.dims.patch_kpad == 592
    res = eng.encode_patches(eng.preprocess_same_size(u8.cuda()), B)
    cos = torch.nn.functional.cosine_similarity(res.cpu().double(), ref.double(), dim=-1)
    rel = ((res.cpu() - ref).norm() / ref.norm()).item()
    print(f"\n[L/14@336 {mode}] cos min {cos.min():.6f}, relative embedding error {rel:.2e}")
    assert cos.min().item() >= 0.999
    assert rel < (3e-3 if mode == "f16" else 2e-2)
    # encode_image entry point (float CHW input) goes through the patchify kernel with the K = 588 -> 592 padding
    emb2 = model.encode_image(x.cuda())
    assert torch.allclose(emb2, res, rtol=1e-3, atol=1e-3)


def test_vit_l14_336_lora_gradients(iic):
    """BASELINE configs[4], training half: one forward + backward of the ViT-L/14@336 tower (T = 577: key blocks in the
    tcgen05 attention forward, column blocks in its backward, rank-16 down-projections on the GEMM) against CPU fp32
    autograd through the oracle, batch 2."""
    import copy
    from oracle import clip_ref, ref_semantics as RS
    om = clip_ref.build_model("ViT-L/14@336px", seed=3)
    sd = {k: v.clone() for k, v in om.state_dict().items()}
    RS.replace_linears_with_lora(om.visual, rank=16, alpha=32)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for n, p in om.visual.named_parameters():
            if n.endswith("lora_A"):
                p.data = torch.randn(p.shape, generator=g) * 0.02
            elif n.endswith("lora_B"):
                p.data = (torch.randn(p.shape, generator=g) * 0.01).to(torch.bfloat16).float()
    for p in om.parameters():
        p.requires_grad_(False)
    lora = {n: p for n, p in om.visual.named_parameters() if "lora" in n and ".mlp." in n}
    for p in lora.values():
        p.requires_grad_(True)
    B = 2
    u8 = torch.randint(0, 256, (B, 336, 336, 3), dtype=torch.uint8, generator=g)
    text = torch.nn.functional.normalize(torch.randn(B, 768, generator=g), dim=-1)
    mean = torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(1, 3, 1, 1)
    std = torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(1, 3, 1, 1)
    x = (u8.permute(0, 3, 1, 2).float() / 255 - mean) / std
    f = om.encode_image(x)
    f = f / f.norm(dim=-1, keepdim=True)
    logits = (f @ text.t()) * 100.0
    labels = torch.arange(B)
    loss_ref = (torch.nn.functional.cross_entropy(logits, labels) + torch.nn.functional.cross_entropy(logits.t(), labels)) / 2
    loss_ref.backward()
    model, _ = iic.load("ViT-L/14@336px", device="cuda", state_dict=sd, operand_dtype="f16")
    iic.replace_linears_with_lora(model.visual, rank=16, alpha=32)
    src = dict(om.visual.named_parameters())
    for n, p in model.visual.named_parameters():
        if "lora" in n:
            p.data = src[n].detach().clone().to(p.device)
    trainer = iic.VisionLoRATrainer(model, logit_scale=100.0)
    loss = trainer.forward_backward(u8.cuda(), text.cuda())
    torch.cuda.synchronize()
    assert abs(loss.item() - loss_ref.item()) < 5e-3 * max(1.0, abs(loss_ref.item())), (loss.item(), loss_ref.item())
    named = dict(model.visual.named_parameters())
    worst = 0.0
    for n, p in lora.items():
        got = named[n].grad
        assert got is not None and got.shape == p.grad.shape, n
        worst = max(worst, ((got.cpu().double() - p.grad.double()).norm() / p.grad.double().norm().clamp_min(1e-30)).item())
    print(f"\n[L/14@336 f16 r=16] loss {loss.item():.5f} (ref {loss_ref.item():.5f}); worst LoRA-gradient relative error {worst:.2e}")
    assert worst < 1e-2
