"""Training path on a B200: backward operators against torch.autograd (fp32) on the same 16-bit operands, and the
LoRA gradients of the whole encoder against the oracle's CPU fp32 autograd (north star: within 1e-2 relative)."""
import copy

import pytest
import torch

from _common import golden_npz, oracle_model, oracle_state_dict

pytestmark = pytest.mark.gpu


def _dump(name, metrics):
    import json, os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    path = os.path.join(out, "train_parity_metrics.json")
    allm = json.load(open(path)) if os.path.exists(path) else {}
    allm[name] = metrics
    json.dump(allm, open(path, "w"), indent=1)


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("B,T,H", [(3, 197, 12), (40, 197, 12), (2, 256, 12), (5, 50, 8), (2, 129, 12), (1, 16, 12), (2, 300, 4), (2, 577, 16), (12, 577, 16)])
@pytest.mark.parametrize("mode", ["bf16", "f16"])
def test_attention_backward(engine, engine_f16, mode, B, T, H):
    """tcgen05 backward (attention_bwd_sm100.cu): one column block for T <= 256, blocks of <= 192 accumulating in TMEM for the
    longer sequences (ViT-L/14 @ 336: T = 577); the larger batches make every SM walk several (image, head) items"""
    eng = engine if mode == "bf16" else engine_f16
    dt = eng.op_dtype
    d = H * 64
    g = torch.Generator(device="cuda").manual_seed(21)
    qkv = (torch.randn(B * T, 3 * d, device="cuda", generator=g)).to(dt)
    do = (torch.randn(B * T, d, device="cuda", generator=g)).to(dt)
    out, dqkv, lse = eng.op_attention_bwd(qkv, do, B, T, H)
    x = qkv.float().clone().requires_grad_(True)
    q, k, v = x.view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(B * T, d)
    ref.backward(do.float())
    assert _rel(out.float(), ref.detach()) < 1e-2
    # lse (log2 domain) of the scaled scores
    s = (q.detach() @ k.detach().transpose(-1, -2)) / 8.0
    lse_ref = torch.logsumexp(s, dim=-1) / 0.6931471805599453          # [B, H, T]
    assert torch.allclose(lse.view(B, H, T), lse_ref, rtol=1e-3, atol=2e-2)
    # P and dS are rounded to the 16-bit operand format before the second product: 2^-8 (bf16) / 2^-11 (fp16) relative
    tol = 1.5e-2 if mode == "bf16" else 4e-3
    gq, gk, gv = x.grad.view(B * T, 3, d).unbind(1)
    dq, dk, dv = dqkv.float().view(B * T, 3, d).unbind(1)
    assert _rel(dq, gq) < tol and _rel(dk, gk) < tol and _rel(dv, gv) < tol, (_rel(dq, gq), _rel(dk, gk), _rel(dv, gv))


@pytest.mark.parametrize("B,T", [(3, 77), (40, 77), (2, 200)])
def test_attention_backward_causal(iic, B, T):
    """the same kernel with the text tower's causal mask (engine created with causal=True): P[q, k] = 0 for k > q"""
    H = 8
    arch = iic.VisionArch(image_size=224, patch_size=16, width=512, layers=1, heads=H, embed_dim=512, seq_tokens=T, causal=True)
    eng = iic.Engine(arch, "cuda:0", operand_dtype="bf16")
    d = H * 64
    g = torch.Generator(device="cuda").manual_seed(27)
    qkv = (torch.randn(B * T, 3 * d, device="cuda", generator=g)).to(torch.bfloat16)
    do = (torch.randn(B * T, d, device="cuda", generator=g)).to(torch.bfloat16)
    out, dqkv, lse = eng.op_attention_bwd(qkv, do, B, T, H)
    x = qkv.float().clone().requires_grad_(True)
    q, k, v = x.view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True).permute(0, 2, 1, 3).reshape(B * T, d)
    ref.backward(do.float())
    assert _rel(out.float(), ref.detach()) < 1e-2
    gq, gk, gv = x.grad.view(B * T, 3, d).unbind(1)
    dq, dk, dv = dqkv.float().view(B * T, 3, d).unbind(1)
    assert _rel(dq, gq) < 1.5e-2 and _rel(dk, gk) < 1.5e-2 and _rel(dv, gv) < 1.5e-2, (_rel(dq, gq), _rel(dk, gk), _rel(dv, gv))


def test_layernorm_and_activation_backward(engine):
    g = torch.Generator(device="cuda").manual_seed(22)
    rows, D = 197 * 3 + 5, 768
    x = (torch.randn(rows, D, device="cuda", generator=g) * 2 + 0.3).requires_grad_(True)
    gamma = torch.randn(D, device="cuda", generator=g)
    beta = torch.randn(D, device="cuda", generator=g)
    dy = torch.randn(rows, D, device="cuda", generator=g).to(torch.bfloat16)
    dx0 = torch.randn(rows, D, device="cuda", generator=g)
    torch.nn.functional.layer_norm(x, (D,), gamma, beta, 1e-5).backward(dy.float())
    dx = dx0.clone()
    dx16 = engine.op_layernorm_bwd(dy, x.detach(), gamma, dx)
    assert torch.allclose(dx, dx0 + x.grad, rtol=1e-4, atol=1e-4), (dx - dx0 - x.grad).abs().max()   # accumulates
    assert torch.equal(dx16, dx.to(torch.bfloat16))
    for act, fn in ((1, lambda u: u * torch.sigmoid(1.702 * u)), (2, torch.nn.functional.gelu)):
        u = torch.randn(1000, 3072, device="cuda", generator=g).to(torch.bfloat16)
        dh = torch.randn(1000, 3072, device="cuda", generator=g).to(torch.bfloat16)
        uu = u.float().requires_grad_(True)
        fn(uu).backward(dh.float())
        got = dh.clone()
        engine.op_act_bwd(got, u, act)
        assert torch.allclose(got.float(), uu.grad, rtol=2 ** -7, atol=1e-3)


@pytest.mark.parametrize("rank", [4, 16])
@pytest.mark.parametrize("shape", [(197 * 9 + 3, 3072), (197 * 128, 3072), (197 * 128, 768), (70, 64)])
def test_lora_gradient_reductions(engine, rank, shape):
    g = torch.Generator(device="cuda").manual_seed(23)
    M, N = shape      # the full-size shapes exercise the whole-wave CTA geometry (rows per CTA depend on M, N and the SM count)
    P = torch.zeros(M, 16, device="cuda", dtype=torch.bfloat16)
    P[:, :rank] = torch.randn(M, rank, device="cuda", generator=g).to(torch.bfloat16)
    Y = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16)
    dB = engine.op_lora_outer(P, Y, rank)                                       # [r, N] = P^T Y
    ref = P[:, :rank].float().t() @ Y.float()
    tol = max(1.0, (M / 1776) ** 0.5)
    assert torch.allclose(dB, ref, rtol=1e-3, atol=2e-2 * tol), (dB - ref).abs().max()
    dA = engine.op_lora_outer(P, Y, rank, act=1, scale=2.0, transpose=True)     # [N, r] = 2 * gelu(Y)^T P
    yf = Y.float()
    # the tensor-core reduction feeds gelu(Y) rounded to the operand type - the same value the forward's c_proj GEMM consumed
    h = (yf * torch.sigmoid(1.702 * yf)).to(torch.bfloat16).float()
    ref = 2.0 * h.t() @ P[:, :rank].float()
    assert torch.allclose(dA, ref, rtol=1e-3, atol=3e-2 * tol), (dA - ref).abs().max()
    assert torch.equal(dA, engine.op_lora_outer(P, Y, rank, act=1, scale=2.0, transpose=True))      # fixed summation order
    ref32 = 2.0 * (yf * torch.sigmoid(1.702 * yf)).t() @ P[:, :rank].float()
    assert (dA - ref32).abs().max() < 1e-2 * ref32.abs().max()      # vs un-rounded gelu: the operand rounding, 2^-9 relative per term


@pytest.mark.parametrize("rank", [4, 16])
@pytest.mark.parametrize("shape", [(197 * 9 + 3, 3072), (197 * 9 + 3, 768), (197 * 128, 3072), (197 * 128, 768), (577 * 2, 1024), (40, 256)])
def test_lora_bwd_fused(engine, rank, shape):
    """dB = P^T . Y and dP = Y . B^T from one pass over Y, against fp32 products of the same bf16 operands"""
    M, N = shape
    g = torch.Generator(device="cuda").manual_seed(29)
    P = torch.zeros(M, 16, device="cuda", dtype=torch.bfloat16)
    P[:, :rank] = torch.randn(M, rank, device="cuda", generator=g).to(torch.bfloat16)
    Bm = torch.zeros(16, N, device="cuda", dtype=torch.bfloat16)
    Bm[:rank] = (torch.randn(rank, N, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    Y = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16)
    db, dp = engine.op_lora_bwd(P, Y, Bm, rank, scale=0.5)
    ref_db = 0.5 * P[:, :rank].float().t() @ Y.float()
    ref_dp = Y.float() @ Bm.float().t()
    assert torch.allclose(db, ref_db, rtol=1e-3, atol=2e-2 * max(1.0, (M / 1776) ** 0.5)), (db - ref_db).abs().max()
    # 16-bit output: 2^-9 relative
    assert torch.allclose(dp.float(), ref_dp, rtol=4e-3, atol=4e-3), (dp.float() - ref_dp).abs().max()
    assert rank == 16 or dp[:, rank:].abs().max() == 0
    db2, dp2 = engine.op_lora_bwd(P, Y, Bm, rank, scale=0.5)     # fixed summation order: bit-reproducible
    assert torch.equal(db, db2) and torch.equal(dp, dp2)


@pytest.mark.parametrize("ctas", [1, 2], ids=["cta1", "cta2"])
@pytest.mark.parametrize("act", [0, 1], ids=["quickgelu", "erf"])
def test_gemm_training_epilogues(iic, engine, ctas, act):
    """c_fc forward with two outputs (activation + kept pre-activation) and the c_proj dX GEMM with act'(u) applied in the
    epilogue, against fp32 references from the same bf16 operands."""
    L = iic._lib
    M, d, mlp = 197 * 5 + 7, 768, 3072
    g = torch.Generator(device="cuda").manual_seed(31)
    y2 = torch.randn(M, d, device="cuda", generator=g).to(torch.bfloat16)
    w1 = (torch.randn(mlp, d, device="cuda", generator=g) * d ** -0.5).to(torch.bfloat16)
    b1 = torch.randn(mlp, device="cuda", generator=g) * 0.2
    p = torch.zeros(M, 16, device="cuda", dtype=torch.bfloat16)
    p[:, :4] = (torch.randn(M, 4, device="cuda", generator=g) * 0.3).to(torch.bfloat16)
    bt = torch.zeros(mlp, 16, device="cuda", dtype=torch.bfloat16)
    bt[:, :4] = (torch.randn(mlp, 4, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    h, u = engine.op_gemm_act_dual(y2, w1, b1, act=act, lora_p=p, lora_bt=bt, r_pad=16, ctas=ctas)
    u_ref = y2.float() @ w1.float().t() + b1 + p.float() @ bt.float().t()
    fwd = (lambda v: v * torch.sigmoid(1.702 * v)) if act == 0 else (lambda v: torch.nn.functional.gelu(v))
    # bf16 rounding of the outputs: 2^-9 relative
    assert torch.allclose(u.float(), u_ref, rtol=4e-3, atol=4e-3), (u.float() - u_ref).abs().max()
    assert torch.allclose(h.float(), fwd(u_ref), rtol=4e-3, atol=4e-3), (h.float() - fwd(u_ref)).abs().max()
    # same pre-activation as the single-output epilogue produces
    assert torch.equal(u, engine.op_gemm(y2, w1, L.EPI_BIAS_BF16, bias=b1, lora_p=p, lora_bt=bt, r_pad=16, ctas=ctas))
    assert torch.equal(h, engine.op_gemm(y2, w1, L.EPI_GELU_ERF_BF16 if act else L.EPI_BIAS_GELU_BF16, bias=b1, lora_p=p, lora_bt=bt,
                                         r_pad=16, ctas=ctas))
    # backward: du = (dy . W2) o act'(u), with W2^T stored [mlp, d] as the engine keeps it
    dy = torch.randn(M, d, device="cuda", generator=g).to(torch.bfloat16)
    w2t = (torch.randn(mlp, d, device="cuda", generator=g) * mlp ** -0.5).to(torch.bfloat16)
    du = engine.op_gemm(dy, w2t, L.EPI_ACT_GRAD_BF16, residual=u, group=2 if act else 1, ctas=ctas,
                        out=torch.empty(M, mlp, device="cuda", dtype=torch.bfloat16))
    uf = u.float().requires_grad_(True)
    fwd(uf).backward(dy.float() @ w2t.float().t())
    assert torch.allclose(du.float(), uf.grad, rtol=6e-3, atol=2e-3), (du.float() - uf.grad).abs().max()


GRAD_BAR = 1e-2          # north star: LoRA gradients within 1e-2 relative - asserted on the default operand dtype
BF16_ARM_GRAD_BAR = 3e-2  # explicit non-default arm: gradients evaluated at activations that carry the 8-bit forward error


@pytest.mark.parametrize("rank", [4, 16])
@pytest.mark.parametrize("mode", ["default", "bf16_arm"])
def test_lora_gradients_match_oracle_autograd(iic, mode, rank):
    """whole encoder, batch 8: d loss / d lora_{A,B} of every vision MLP vs CPU fp32 autograd through the oracle
    (rank 4: main.py's adapters, down-projections fused into LayerNorm / the c_fc epilogue; rank 16: train_lora.py's,
    down-projections on the tcgen05 GEMM)"""
    from oracle import ref_semantics as RS
    B = 8
    crops = torch.from_numpy(golden_npz("crops_u8.npz")["crops"][:B])
    text = torch.from_numpy(golden_npz("text_features.npz")["text"][40:40 + B]).clone()
    # ---- oracle: reference LoRA wrap (main.py:62-74) + train_lora.py's loss, fp32 CPU autograd ----
    om = copy.deepcopy(oracle_model())
    RS.replace_linears_with_lora(om, rank=rank, alpha=2 * rank)
    RS.seed_vision_lora(om, seed=99)
    for p in om.parameters():
        p.requires_grad_(False)
    lora = {n: p for n, p in om.named_parameters() if n.startswith("visual.") and "lora" in n and ".mlp." in n}
    for p in lora.values():
        p.requires_grad_(True)
    mean = torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(1, 3, 1, 1)
    std = torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(1, 3, 1, 1)
    x = (crops.permute(0, 3, 1, 2).float() / 255 - mean) / std
    f = om.encode_image(x)
    f = f / f.norm(dim=-1, keepdim=True)
    logits = (f @ text.t()) * 100.0
    labels = torch.arange(B)
    loss_ref = (torch.nn.functional.cross_entropy(logits, labels) + torch.nn.functional.cross_entropy(logits.t(), labels)) / 2
    loss_ref.backward()
    # ---- product ----
    model, _ = iic.load("ViT-B/16", device="cuda", state_dict=oracle_state_dict(),
                        operand_dtype=None if mode == "default" else "bf16")
    iic.replace_linears_with_lora(model, rank=rank, alpha=2 * rank)
    src = {n: p for n, p in om.named_parameters() if "lora" in n}
    for n, p in model.named_parameters():
        if n in src:
            p.data = src[n].detach().clone().to(p.device)
    trainer = iic.VisionLoRATrainer(model, logit_scale=100.0)
    loss = trainer.forward_backward(crops.cuda(), text.cuda())
    torch.cuda.synchronize()
    assert abs(loss.item() - loss_ref.item()) < 5e-3 * max(1.0, abs(loss_ref.item())), (loss.item(), loss_ref.item())
    named = dict(model.named_parameters())
    errs = {}
    for n, p in lora.items():
        got = named[n].grad
        assert got is not None and got.shape == p.grad.shape, n
        errs[n] = _rel(got.cpu(), p.grad)
    worst = max(errs.values())
    by_layer = [max(v for k, v in errs.items() if f"resblocks.{i}." in k) for i in range(12)]
    print(f"\n[{mode} r={rank}] loss {loss.item():.5f} (ref {loss_ref.item():.5f}); worst LoRA-gradient relative error {worst:.2e} over "
          f"{len(lora)} tensors; per block: " + " ".join(f"{e:.1e}" for e in by_layer))
    _dump(f"{'f16' if mode == 'default' else 'bf16'}_r{rank}", {"loss": loss.item(), "loss_ref": loss_ref.item(), "worst_rel": worst, "per_block_worst_rel": by_layer})
    # north star: LoRA gradients within 1e-2 relative, asserted on the default dtype (fp16 operands + power-of-two loss
    # scaling; measured 2e-3).  The explicit bf16 arm evaluates the gradient at activations that already carry the
    # 8-bit-mantissa forward error (every block sits at ~1e-2: not an accumulation effect; measured worst 1.3e-2).
    assert worst < (GRAD_BAR if mode == "default" else BF16_ARM_GRAD_BAR), max(errs.items(), key=lambda kv: kv[1])
    # out_proj LoRA parameters get no gradient (dead in the reference's forward, F4)
    assert all(named[n].grad is None for n in named if ".attn.out_proj.lora." in n)


def _text_tokens(B, seed):
    g = torch.Generator().manual_seed(seed)
    tok = torch.zeros(B, 77, dtype=torch.long)
    for b in range(B):
        n = int(torch.randint(4, 76, (1,), generator=g))
        tok[b, 0] = 49406
        tok[b, 1:n] = torch.randint(1, 49405, (n - 1,), generator=g)
        tok[b, n] = 49407                       # EOT = the largest id: encode_text takes the row at argmax
    return tok


@pytest.mark.parametrize("rank", [4, 16])
@pytest.mark.parametrize("mode", ["default", "bf16_arm"])
def test_text_lora_gradients_match_oracle_autograd(iic, mode, rank):
    """train_lora.py's own step (text tower through LoRA, image features fixed): d loss / d lora_{A,B} of every text MLP
    vs CPU fp32 autograd through the oracle - causal attention backward, sequence forward / backward of the engine"""
    from oracle import ref_semantics as RS
    B = 8
    tokens = _text_tokens(B, seed=5)
    g = torch.Generator().manual_seed(6)
    img = torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=-1)
    om = copy.deepcopy(oracle_model())
    RS.replace_linears_with_lora(om, rank=rank, alpha=2 * rank)
    gen = torch.Generator().manual_seed(77)
    for n, p in om.named_parameters():
        p.requires_grad_(False)
        if n.startswith("transformer.") and n.endswith("lora_A"):
            p.data = torch.randn(p.shape, generator=gen) * 0.02
        if n.startswith("transformer.") and n.endswith("lora_B"):
            p.data = torch.randn(p.shape, generator=gen) * 0.004
    lora = {n: p for n, p in om.named_parameters() if n.startswith("transformer.") and "lora" in n and ".mlp." in n}
    for p in lora.values():
        p.requires_grad_(True)
    f = om.encode_text(tokens)
    f = f / f.norm(dim=-1, keepdim=True)
    logits = (img @ f.t()) * 100.0
    labels = torch.arange(B)
    loss_ref = (torch.nn.functional.cross_entropy(logits, labels) + torch.nn.functional.cross_entropy(logits.t(), labels)) / 2
    loss_ref.backward()
    model, _ = iic.load("ViT-B/16", device="cuda", state_dict=oracle_state_dict(),
                        operand_dtype=None if mode == "default" else "bf16")
    iic.replace_linears_with_lora(model, rank=rank, alpha=2 * rank)
    src = {n: p for n, p in om.named_parameters() if "lora" in n}
    for n, p in model.named_parameters():
        if n in src:
            p.data = src[n].detach().clone().to(p.device)
    trainer = iic.TextLoRATrainer(model, logit_scale=100.0)
    loss = trainer.forward_backward(img.cuda(), tokens.cuda())
    torch.cuda.synchronize()
    assert abs(loss.item() - loss_ref.item()) < 5e-3 * max(1.0, abs(loss_ref.item())), (loss.item(), loss_ref.item())
    named = dict(model.named_parameters())
    errs = {n: _rel(named[n].grad.cpu(), p.grad) for n, p in lora.items()}
    worst = max(errs.values())
    print(f"\n[text {mode} r={rank}] loss {loss.item():.5f} (ref {loss_ref.item():.5f}); worst text-LoRA gradient relative error {worst:.2e} "
          f"over {len(lora)} tensors")
    _dump(f"text_{'f16' if mode == 'default' else 'bf16'}_r{rank}", {"loss": loss.item(), "loss_ref": loss_ref.item(), "worst_rel": worst})
    # the default dtype is held to the north star's 1e-2; the explicit bf16 arm carries the 8-bit-mantissa forward error into
    # every block's gradient (vision tower: 1.3e-2; the 77-token text tower with its near-uniform initial loss: 2.0e-2)
    assert worst < (GRAD_BAR if mode == "default" else BF16_ARM_GRAD_BAR), max(errs.items(), key=lambda kv: kv[1])
    # a second call reproduces the gradients bit for bit, and the inference path still agrees with the trained parameters
    g1 = {n: named[n].grad.clone() for n in lora}
    trainer.forward_backward(img.cuda(), tokens.cuda())
    assert all(torch.equal(g1[n], named[n].grad) for n in lora)


def test_text_training_step_reduces_loss(iic):
    """a few steps of the reference's real training step (text-side LoRA) lower the loss and move only text LoRA parameters"""
    B = 8
    tokens = _text_tokens(B, seed=8).cuda()
    g = torch.Generator().manual_seed(9)
    img = torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=-1).cuda()
    model, _ = iic.load("ViT-B/16", device="cuda", state_dict=oracle_state_dict())
    iic.replace_linears_with_lora(model, rank=16, alpha=32)
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    trainer = iic.TextLoRATrainer(model, lr=1e-3)
    losses = [trainer.step(img, tokens) for _ in range(8)]
    assert losses[-1] < losses[0] - 0.05, losses
    moved = {n for n, p in model.named_parameters() if not torch.equal(p.detach(), before[n])}
    assert moved and all(n.startswith("transformer.") and ".mlp." in n and "lora" in n for n in moved), sorted(moved)[:5]
    # model.encode_text (inference path on the engine) sees the trained adapters
    model.text_on_engine = True
    with torch.no_grad():
        f = model.encode_text(tokens).float()
    f = f / f.norm(dim=-1, keepdim=True)
    logits = (img @ f.t()) * float(model.logit_scale.exp())
    labels = torch.arange(B, device="cuda")
    l_inf = (torch.nn.functional.cross_entropy(logits, labels) + torch.nn.functional.cross_entropy(logits.t(), labels)) / 2
    assert abs(float(l_inf) - trainer.step(img, tokens)) < 0.05


def test_training_step_reduces_loss(iic):
    """a few AdamW steps (train_lora.py:249-252) on one batch lower the loss and move only LoRA parameters"""
    B = 8
    crops = torch.from_numpy(golden_npz("crops_u8.npz")["crops"][:B]).cuda()
    text = torch.from_numpy(golden_npz("text_features.npz")["text"][40:40 + B]).cuda()
    model, _ = iic.load("ViT-B/16", device="cuda", state_dict=oracle_state_dict())
    iic.replace_linears_with_lora(model, rank=4, alpha=8)            # fresh LoRA: lora_B == 0 (main.py:27)
    frozen = {n: p.detach().clone() for n, p in model.named_parameters() if "lora" not in n and n.startswith("visual.")}
    trainer = iic.VisionLoRATrainer(model, lr=1e-3, logit_scale=100.0)
    losses = [trainer.step(crops, text) for _ in range(6)]
    assert losses[-1] < losses[0], losses
    for n, p in model.named_parameters():
        if n in frozen:
            assert torch.equal(p.detach(), frozen[n]), n
    assert any(bool((p != 0).any()) for n, p in model.named_parameters() if n.endswith("mlp.c_fc.lora.lora_B") and n.startswith("visual."))


def test_training_step_cuda_graph_matches_eager(iic):
    """VisionLoRATrainer(use_graph=True): forward + loss + backward replayed as one CUDA graph gives the gradients of the eager
    launches (same kernels; only the power-of-two loss scale may differ), keeps doing so as the parameters move under the
    optimizer (the per-step LoRA operand refresh is inside the graph), and the loss goes down."""
    B = 8
    crops = torch.from_numpy(golden_npz("crops_u8.npz")["crops"][:B]).cuda()
    text = torch.from_numpy(golden_npz("text_features.npz")["text"][40:40 + B]).cuda()

    def make():
        model, _ = iic.load("ViT-B/16", device="cuda", state_dict=oracle_state_dict())
        iic.replace_linears_with_lora(model, rank=4, alpha=8)
        g = torch.Generator().manual_seed(3)
        for n, p in model.named_parameters():
            if n.startswith("visual.") and ".mlp." in n and n.endswith("lora_B"):
                p.data = (torch.randn(p.shape, generator=g) * 0.004).cuda()
        return model
    eager = iic.VisionLoRATrainer(make(), lr=1e-3, logit_scale=100.0)
    graphed = iic.VisionLoRATrainer(make(), lr=1e-3, logit_scale=100.0, use_graph=True)
    le, lg = [], []
    for step in range(5):
        le.append(eager.step(crops, text))
        lg.append(graphed.step(crops, text))
        ge, gg = eager.flat_grads, graphed.flat_grads
        rel = ((ge - gg).norm() / ge.norm()).item()
        assert rel < 2e-3, (step, rel)
        assert abs(le[-1] - lg[-1]) < 1e-3 * max(1.0, abs(le[-1])), (le, lg)
    assert graphed._graph is not None and graphed._graph_steps >= 4      # replays, not re-captures
    assert lg[-1] < lg[0]
    pe = torch.cat([p.detach().reshape(-1) for p in eager.params])
    pg = torch.cat([p.detach().reshape(-1) for p in graphed.params])
    assert ((pe - pg).norm() / pe.norm()).item() < 1e-3
