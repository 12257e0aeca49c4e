"""Per-kernel numerics on a B200, through the C ABI (`iic_op_*`), against plain PyTorch fp32 references of the same
op computed from the SAME bf16-rounded operands (so the only differences are accumulation order and the final
rounding).  Tolerances are written next to each check.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

L = None


@pytest.fixture(scope="module", autouse=True)
def _consts(iic):
    global L
    L = iic._lib


def _bf16(t):
    return t.to(torch.bfloat16)


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def quick_gelu(x):
    return x * torch.sigmoid(1.702 * x)


GEMM_SHAPES = [
    # (M, N, K): one full tile; ragged M; multi-tile persistent; encoder shapes at a small batch
    (128, 256, 64),
    (256, 256, 768),
    (197 * 3, 768, 768),
    (197 * 16, 2304, 768),
    (197 * 16, 3072, 768),
    (197 * 16, 768, 3072),
    (197 * 40 + 5, 768, 768),
]


@pytest.mark.parametrize("ctas", [1, 2], ids=["cta1", "cta2"])
@pytest.mark.parametrize("shape", GEMM_SHAPES)
def test_gemm_bias_bf16(engine, shape, ctas):
    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N + K)
    a = _bf16(torch.randn(M, K, device="cuda", generator=g))
    w = _bf16(torch.randn(N, K, device="cuda", generator=g) * K ** -0.5)
    bias = torch.randn(N, device="cuda", generator=g)
    out = engine.op_gemm(a, w, L.EPI_BIAS_BF16, bias=bias, ctas=ctas)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias
    # bf16 output rounding: 2^-9 relative per element; fp32 accumulation order differences are far below that
    assert torch.allclose(out.float(), ref, rtol=2 ** -7, atol=2e-2), (out.float() - ref).abs().max()
    assert _rel(out.float(), ref) < 4e-3


@pytest.mark.parametrize("ctas", [1, 2], ids=["cta1", "cta2"])
def test_gemm_quickgelu(engine, ctas):
    M, N, K = 197 * 8, 3072, 768
    g = torch.Generator(device="cuda").manual_seed(1)
    a = _bf16(torch.randn(M, K, device="cuda", generator=g))
    w = _bf16(torch.randn(N, K, device="cuda", generator=g) * K ** -0.5)
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    out = engine.op_gemm(a, w, L.EPI_BIAS_GELU_BF16, bias=bias, ctas=ctas)
    ref = quick_gelu(a.float() @ w.float().t() + bias)
    assert torch.allclose(out.float(), ref, rtol=2 ** -7, atol=1e-2), (out.float() - ref).abs().max()
    out2 = engine.op_gemm(a, w, L.EPI_GELU_ERF_BF16, bias=bias, ctas=ctas)
    ref2 = torch.nn.functional.gelu(a.float() @ w.float().t() + bias)
    assert torch.allclose(out2.float(), ref2, rtol=2 ** -7, atol=1e-2), (out2.float() - ref2).abs().max()


@pytest.mark.parametrize("ctas", [1, 2], ids=["cta1", "cta2"])
def test_gemm_residual_f32_inplace(engine, ctas):
    M, N, K = 197 * 8 + 3, 768, 3072
    g = torch.Generator(device="cuda").manual_seed(2)
    a = _bf16(torch.randn(M, K, device="cuda", generator=g))
    w = _bf16(torch.randn(N, K, device="cuda", generator=g) * K ** -0.5)
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    x = torch.randn(M, N, device="cuda", generator=g)
    ref = x + a.float() @ w.float().t() + bias
    engine.op_gemm(a, w, L.EPI_BIAS_RES_F32, bias=bias, residual=x, out=x, ctas=ctas)  # in place on the stream
    # fp32 in, fp32 out: only the accumulation order differs
    assert torch.allclose(x, ref, rtol=1e-4, atol=2e-4), (x - ref).abs().max()


@pytest.mark.parametrize("ctas", [1, 2], ids=["cta1", "cta2"])
def test_gemm_patch_embed_scatter(engine, ctas):
    B, G, N, K = 5, 196, 768, 768
    g = torch.Generator(device="cuda").manual_seed(3)
    a = _bf16(torch.randn(B * G, K, device="cuda", generator=g))
    w = _bf16(torch.randn(N, K, device="cuda", generator=g) * K ** -0.5)
    pos = torch.randn(G + 1, N, device="cuda", generator=g)
    out = torch.full((B * (G + 1), N), 7.0, device="cuda")
    engine.op_gemm(a, w, L.EPI_POS_F32, residual=pos, out=out, group=G, ctas=ctas)
    ref = (a.float() @ w.float().t()).view(B, G, N) + pos[1:]
    got = out.view(B, G + 1, N)
    assert torch.allclose(got[:, 1:], ref, rtol=1e-4, atol=2e-4)
    assert (got[:, 0] == 7.0).all()  # class-token rows are not touched by the GEMM


@pytest.mark.parametrize("ctas", [1, 2], ids=["cta1", "cta2"])
@pytest.mark.parametrize("rank", [4, 16])
def test_gemm_lora_fused(engine, ctas, rank):
    """reference semantic: linear(x) + (x @ A @ B) * scaling  (/root/reference/main.py:30-31, 42-43)"""
    M, N, K = 197 * 6, 3072, 768
    g = torch.Generator(device="cuda").manual_seed(4 + rank)
    x = _bf16(torch.randn(M, K, device="cuda", generator=g))
    w = _bf16(torch.randn(N, K, device="cuda", generator=g) * K ** -0.5)
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    A = torch.randn(K, rank, device="cuda", generator=g) * 0.02
    Bm = torch.randn(rank, N, device="cuda", generator=g) * 0.3   # large on purpose: the delta must be visible
    scaling = 2.0
    r4 = (rank + 3) // 4 * 4
    a_scaled = torch.zeros(K, r4, device="cuda")
    a_scaled[:, :rank] = A * scaling
    p = engine.op_lora_down(x, a_scaled)
    p_ref = (x.float() @ A) * scaling
    assert torch.allclose(p[:, :rank].float(), p_ref, rtol=2 ** -7, atol=1e-3)
    assert (p[:, rank:] == 0).all()
    bt = torch.zeros(N, 16, device="cuda", dtype=torch.bfloat16)
    bt[:, :rank] = _bf16(Bm.t())
    out = engine.op_gemm(x, w, L.EPI_BIAS_BF16, bias=bias, lora_p=p, lora_bt=bt, r_pad=16, ctas=ctas)
    base = x.float() @ w.float().t() + bias
    ref = base + p[:, :rank].float() @ bt[:, :rank].float().t()
    assert torch.allclose(out.float(), ref, rtol=2 ** -7, atol=2e-2), (out.float() - ref).abs().max()
    # and the delta is really there (a skipped LoRA block would fail this)
    assert (out.float() - base).abs().mean() > 10 * (out.float() - ref).abs().mean()
    # against the un-rounded reference formula
    full = base + (x.float() @ A @ Bm) * scaling
    assert _rel(out.float(), full) < 6e-3


# rows of 1 / 2 / 4 / 8 images: the launcher cuts these into 64- or 128-column tiles (gemm_tile_n) so that the single-image path
# does not leave most SMs idle; the last one is wide again
NARROW_M = [197, 2 * 197, 4 * 197, 8 * 197, 77, 33]


@pytest.mark.parametrize("M", NARROW_M)
def test_gemm_narrow_tiles(engine, M):
    """every inference epilogue at small row counts (narrow output tiles): same fp32 references and tolerances as the wide-tile tests"""
    g = torch.Generator(device="cuda").manual_seed(100 + M)
    d, mlp = 768, 3072
    x = _bf16(torch.randn(M, d, device="cuda", generator=g))
    for N, K in ((3 * d, d), (d, d), (mlp, d)):
        w = _bf16(torch.randn(N, K, device="cuda", generator=g) * K ** -0.5)
        bias = torch.randn(N, device="cuda", generator=g) * 0.1
        a = x if K == d else None
        ref = a.float() @ w.float().t() + bias
        out = engine.op_gemm(a, w, L.EPI_BIAS_BF16, bias=bias)
        assert torch.allclose(out.float(), ref, rtol=2 ** -7, atol=2e-2), (N, K, (out.float() - ref).abs().max())
        out = engine.op_gemm(a, w, L.EPI_BIAS_GELU_BF16, bias=bias)
        assert torch.allclose(out.float(), quick_gelu(ref), rtol=2 ** -7, atol=1e-2), (N, K)
        out = engine.op_gemm(a, w, L.EPI_GELU_ERF_BF16, bias=bias)
        assert torch.allclose(out.float(), torch.nn.functional.gelu(ref), rtol=2 ** -7, atol=1e-2), (N, K)
    # residual epilogues, in place on the fp32 stream: short and long reduction (the latter takes the deep operand ring)
    for K in (d, mlp):
        a = _bf16(torch.randn(M, K, device="cuda", generator=g))
        w = _bf16(torch.randn(d, K, device="cuda", generator=g) * K ** -0.5)
        bias = torch.randn(d, device="cuda", generator=g) * 0.1
        xs = torch.randn(M, d, device="cuda", generator=g)
        ref = xs + a.float() @ w.float().t() + bias
        engine.op_gemm(a, w, L.EPI_BIAS_RES_F32, bias=bias, residual=xs, out=xs)
        assert torch.allclose(xs, ref, rtol=1e-4, atol=2e-4), (K, (xs - ref).abs().max())
    # LoRA k-step into the same accumulator
    rank = 4
    w = _bf16(torch.randn(mlp, d, device="cuda", generator=g) * d ** -0.5)
    bias = torch.randn(mlp, device="cuda", generator=g) * 0.1
    p = torch.zeros(M, 16, device="cuda", dtype=torch.bfloat16)
    p[:, :rank] = _bf16(torch.randn(M, rank, device="cuda", generator=g))
    bt = torch.zeros(mlp, 16, device="cuda", dtype=torch.bfloat16)
    bt[:, :rank] = _bf16(torch.randn(mlp, rank, device="cuda", generator=g) * 0.3)
    out = engine.op_gemm(x, w, L.EPI_BIAS_GELU_BF16, bias=bias, lora_p=p, lora_bt=bt, r_pad=16)
    base = x.float() @ w.float().t() + bias
    ref = quick_gelu(base + p.float() @ bt.float().t())
    assert torch.allclose(out.float(), ref, rtol=2 ** -7, atol=2e-2), (out.float() - ref).abs().max()
    assert (out.float() - quick_gelu(base)).abs().mean() > 10 * (out.float() - ref).abs().mean()


@pytest.mark.parametrize("B", [1, 2, 4])
def test_gemm_narrow_patch_embed(engine, B):
    G, N, K = 196, 768, 768
    g = torch.Generator(device="cuda").manual_seed(3 + B)
    a = _bf16(torch.randn(B * G, K, device="cuda", generator=g))
    w = _bf16(torch.randn(N, K, device="cuda", generator=g) * K ** -0.5)
    pos = torch.randn(G + 1, N, device="cuda", generator=g)
    out = torch.full((B * (G + 1), N), 7.0, device="cuda")
    engine.op_gemm(a, w, L.EPI_POS_F32, residual=pos, out=out, group=G)
    ref = (a.float() @ w.float().t()).view(B, G, N) + pos[1:]
    got = out.view(B, G + 1, N)
    assert torch.allclose(got[:, 1:], ref, rtol=1e-4, atol=2e-4)
    assert (got[:, 0] == 7.0).all()


def test_small_batch_lora_matches_large_batch(engine):
    """a LoRA-adapted model (rank 4 on c_fc / c_proj: the c_proj down-projection rides in the c_fc epilogue as per-tile partials) gives
    the same embedding for an image alone (narrow tiles, more partial slots) and inside a batch of 40 (wide tiles): the frozen
    products accumulate in the same order, only the grouping of the partials differs (fp32, far below the 16-bit operand rounding)"""
    from importlib import import_module
    clipc = import_module("ai-interior-image-classifier_b200.clip_compat")
    lora = import_module("ai-interior-image-classifier_b200.lora")
    vis = clipc.build_visual("ViT-B/16", seed=0).cuda()
    for blk in vis.transformer.resblocks:
        blk.mlp.c_fc = lora.LoRALinear(blk.mlp.c_fc, rank=4, alpha=8)
        blk.mlp.c_proj = lora.LoRALinear(blk.mlp.c_proj, rank=4, alpha=8)
    gen = torch.Generator().manual_seed(5)
    for n, p_ in vis.named_parameters():
        if n.endswith("lora_B"):
            p_.data.copy_(torch.randn(p_.shape, generator=gen) * 0.02)
    eng = vis.sync_engine()
    text = torch.nn.functional.normalize(torch.randn(60, 512, generator=gen), dim=-1).cuda()
    eng.set_labels(text, [40, 20], [11, 0], topk=5, logit_scale=100.0)
    imgs = torch.randint(0, 256, (40, 224, 224, 3), dtype=torch.uint8, generator=gen).cuda()
    big = eng.classify_same_size(imgs)
    for B in (1, 2, 4):
        small = eng.classify_same_size(imgs[:B].clone(), use_graph=False)
        err = (small.embedding - big.embedding[:B]).abs().max().item()
        assert err <= 2e-3 * big.embedding.abs().max().item(), (B, err)
        assert (small.logits - big.logits[:B]).abs().max().item() < 2e-3



@pytest.mark.parametrize("D", [768, 1024])
def test_layernorm(engine, D):
    rows = 197 * 4 + 1
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(rows, D, device="cuda", generator=g) * 3 + 0.5
    gamma = torch.randn(D, device="cuda", generator=g)
    beta = torch.randn(D, device="cuda", generator=g)
    ref = torch.nn.functional.layer_norm(x, (D,), gamma, beta, 1e-5)
    out32 = engine.op_layernorm(x, gamma, beta, out_dtype=torch.float32)
    assert torch.allclose(out32, ref, rtol=1e-5, atol=1e-5), (out32 - ref).abs().max()
    out16 = engine.op_layernorm(x, gamma, beta, out_dtype=torch.bfloat16)
    assert torch.allclose(out16.float(), ref, rtol=2 ** -8, atol=1e-6)
    A = torch.randn(D, 4, device="cuda", generator=g) * 0.04
    out16b, p = engine.op_layernorm(x, gamma, beta, out_dtype=torch.bfloat16, lora_a_scaled=A)
    assert torch.equal(out16b, out16)
    assert torch.allclose(p[:, :4].float(), ref @ A, rtol=2 ** -7, atol=1e-3)
    assert (p[:, 4:] == 0).all()
    A8 = torch.randn(D, 8, device="cuda", generator=g) * 0.04          # generic-rank path
    _, p8 = engine.op_layernorm(x, gamma, beta, out_dtype=torch.bfloat16, lora_a_scaled=A8)
    assert torch.allclose(p8[:, :8].float(), ref @ A8, rtol=2 ** -7, atol=1e-3) and (p8[:, 8:] == 0).all()


@pytest.mark.parametrize("impl", [1, 2], ids=["mma_sync", "tcgen05"])
@pytest.mark.parametrize("T,B,H", [(197, 3, 12), (577, 2, 16), (50, 2, 12), (16, 1, 12), (197, 40, 12), (256, 2, 12), (129, 1, 12), (257, 3, 12), (577, 20, 16), (640, 1, 4), (1, 2, 12)])
def test_attention(engine, T, B, H, impl):
    d = H * 64
    g = torch.Generator(device="cuda").manual_seed(6)
    qkv = _bf16(torch.randn(B * T, 3 * d, device="cuda", generator=g))
    out = engine.op_attention(qkv, B, T, H, impl=impl)
    q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v)  # fp32 math
    ref = ref.permute(0, 2, 1, 3).reshape(B * T, d)
    # P is rounded to bf16 before the PV product and the output is bf16: 2^-8 relative on values O(1)
    assert torch.allclose(out.float(), ref, rtol=2 ** -6, atol=2e-2), (out.float() - ref).abs().max()
    assert _rel(out.float(), ref) < 8e-3


@pytest.mark.parametrize("mode", ["bf16", "f16"])
@pytest.mark.parametrize("T,B,H", [(197, 3, 12), (197, 40, 12), (197, 1024, 12), (193, 2, 12), (192, 6, 12), (129, 5, 12), (128, 7, 12), (50, 2, 12),
                                   (16, 1, 12), (1, 2, 12), (77, 9, 8), (208, 3, 4), (145, 13, 12), (177, 150, 12)])
def test_attention_whole_row(engine, engine_f16, mode, T, B, H):
    """whole-row tcgen05 kernel (attention_row_sm100.cu, impl 3; the encoder's default for unmasked T <= 208): one / two query
    tiles, odd and even unit counts, padding keys in the last unit, single-unit rows, batches that make every SM walk many
    items; against fp32 SDPA on the same 16-bit operands.  The large batch is the BASELINE configs[2] shape."""
    eng = engine if mode == "bf16" else engine_f16
    d = H * 64
    g = torch.Generator(device="cuda").manual_seed(60 + T)
    qkv = (torch.randn(B * T, 3 * d, device="cuda", generator=g) * 1.5).to(eng.op_dtype)
    out = eng.op_attention(qkv, B, T, H, impl=3)
    ref = torch.empty(B * T, d, device="cuda")
    for b0 in range(0, B, 64):      # chunked: the fp32 reference of 1024 images would need 6 GB of scores
        b1 = min(B, b0 + 64)
        q, k, v = qkv[b0 * T:b1 * T].float().view(b1 - b0, T, 3, H, 64).permute(2, 0, 3, 1, 4)
        ref[b0 * T:b1 * T] = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape((b1 - b0) * T, d)
    assert torch.isfinite(out.float()).all()
    assert torch.allclose(out.float(), ref, rtol=2 ** -6, atol=2e-2), (out.float() - ref).abs().max()
    assert _rel(out.float(), ref) < (8e-3 if mode == "bf16" else 1.5e-3)
    # a second launch gives the same bits, and the default dispatch (impl 0) picks this kernel for unmasked T <= 208
    assert torch.equal(out, eng.op_attention(qkv, B, T, H, impl=3)) and torch.equal(out, eng.op_attention(qkv, B, T, H, impl=0))


def test_attention_whole_row_extreme_scores(engine_f16):
    """rows whose scores span hundreds of units (exact maximum: no rescale path to get wrong), rows that are constant, and one
    dominant key per row: p must stay finite and normalised"""
    eng = engine_f16
    B, T, H = 2, 197, 12
    d = H * 64
    g = torch.Generator(device="cuda").manual_seed(16)
    qkv = torch.randn(B, T, 3, H, 64, device="cuda", generator=g)
    ramp = torch.arange(T, device="cuda", dtype=torch.float32)
    qkv[:, :, 0, :, 0] = 4.0
    qkv[:, :, 1, 0::3, 0] = (0.5 * ramp)[None, :, None]                    # steadily rising scores (max at the last key)
    qkv[:, :, 1, 1::3, 0] = (-0.5 * ramp)[None, :, None]                   # falling: max at key 0
    qkv[:, :, 1, 2::3, :] = 0.0                                            # all scores equal: uniform attention
    qkv = qkv.reshape(B * T, 3 * d).to(torch.float16)
    out = eng.op_attention(qkv, B, T, H, impl=3)
    q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(B * T, d)
    assert torch.isfinite(out.float()).all()
    assert torch.allclose(out.float(), ref, rtol=2 ** -8, atol=4e-3), (out.float() - ref).abs().max()


@pytest.mark.parametrize("T,B,H", [(197, 2, 12), (577, 1, 16)])
def test_attention_rising_scores(engine, T, B, H):
    """Scores that grow along the key axis force the tcgen05 kernel's running maximum to move in every key block
    (the O / row-sum rescale path); a flat tail checks that the lazy maximum (left alone below 2^8) stays exact."""
    d = H * 64
    g = torch.Generator(device="cuda").manual_seed(16)
    qkv = torch.randn(B, T, 3, H, 64, device="cuda", generator=g)
    ramp = torch.arange(T, device="cuda", dtype=torch.float32)
    qkv[:, :, 0, :, 0] = 4.0                                     # q[..., 0]
    qkv[:, :, 1, 0::2, 0] = (0.2 * ramp)[None, :, None]          # even heads: steadily rising scores
    qkv[:, :, 1, 1::2, 0] = (0.05 * ramp.clamp(max=T // 2))[None, :, None]   # odd heads: slow rise, then flat
    qkv = _bf16(qkv.reshape(B * T, 3 * d))
    out = engine.op_attention(qkv, B, T, H, impl=2)
    q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(B * T, d)
    assert torch.allclose(out.float(), ref, rtol=2 ** -6, atol=2e-2), (out.float() - ref).abs().max()
    assert _rel(out.float(), ref) < 8e-3


def test_head_from_embeddings(engine):
    E, groups, split = 512, [40, 20, 12, 299, 36, 30], [11, 0, 0, 0, 0, 0]
    Lc = sum(groups)
    g = torch.Generator(device="cuda").manual_seed(7)
    text = torch.nn.functional.normalize(torch.randn(Lc, E, device="cuda", generator=g), dim=-1)
    engine.set_labels(text, groups, split, topk=5, logit_scale=100.0)
    for B in (1, 4, 7):
        emb = torch.randn(B, E, device="cuda", generator=g) * 3
        r = engine.head(emb)
        f = emb / emb.norm(dim=-1, keepdim=True)
        logits = 100.0 * f @ text.t()
        assert torch.allclose(r.logits, logits, rtol=1e-5, atol=2e-4), (r.logits - logits).abs().max()
        off = 0
        for gi, n in enumerate(groups):
            p = logits[:, off:off + n].softmax(dim=-1)
            assert torch.allclose(r.probs[:, off:off + n], p, rtol=1e-4, atol=1e-6)
            vals, inds = p.topk(min(5, n), dim=-1)
            assert torch.allclose(r.topk_val[:, gi], vals, rtol=1e-4, atol=1e-6)
            assert torch.equal(r.topk_idx[:, gi].long(), inds)
            if split[gi]:
                assert torch.allclose(r.split_sum[:, gi], p[:, :split[gi]].sum(-1), rtol=1e-4, atol=1e-6)
            off += n


# ---------------------------------------------------------------------------------------------- fp16 operand mode
@pytest.mark.parametrize("ctas", [1, 2], ids=["cta1", "cta2"])
def test_f16_gemm_lora_gelu(engine_f16, ctas):
    """same kernels instantiated for fp16 operands (instruction descriptor, TMA dtype, packers)"""
    eng = engine_f16
    M, N, K, rank = 197 * 5 + 7, 3072, 768, 4
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(M, K, device="cuda", generator=g).half()
    w = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).half()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    A = torch.randn(K, rank, device="cuda", generator=g) * 0.04
    Bm = (torch.randn(rank, N, device="cuda", generator=g) * 0.3)
    p = eng.op_lora_down(x, A.contiguous())
    assert p.dtype == torch.float16
    bt = torch.zeros(N, 16, device="cuda", dtype=torch.float16)
    bt[:, :rank] = Bm.t().half()
    out = eng.op_gemm(x, w, L.EPI_BIAS_GELU_BF16, bias=bias, lora_p=p, lora_bt=bt, r_pad=16, ctas=ctas)
    assert out.dtype == torch.float16
    ref = quick_gelu(x.float() @ w.float().t() + bias + p[:, :rank].float() @ bt[:, :rank].float().t())
    # fp16 output rounding: 2^-11 relative
    assert torch.allclose(out.float(), ref, rtol=2 ** -9, atol=3e-3), (out.float() - ref).abs().max()
    x32 = torch.randn(M, 768, device="cuda", generator=g)
    w2 = (torch.randn(768, N, device="cuda", generator=g) * N ** -0.5).half()
    ref2 = x32 + out.float() @ w2.float().t()
    eng.op_gemm(out, w2, L.EPI_BIAS_RES_F32, residual=x32, out=x32, ctas=ctas)
    assert torch.allclose(x32, ref2, rtol=1e-4, atol=3e-4), (x32 - ref2).abs().max()


def test_f16_layernorm_attention(engine_f16):
    eng = engine_f16
    g = torch.Generator(device="cuda").manual_seed(12)
    x = torch.randn(197 * 2, 768, device="cuda", generator=g) * 2
    gamma, beta = torch.randn(768, device="cuda", generator=g), torch.randn(768, device="cuda", generator=g)
    ref = torch.nn.functional.layer_norm(x, (768,), gamma, beta, 1e-5)
    out = eng.op_layernorm(x, gamma, beta)
    assert out.dtype == torch.float16 and torch.allclose(out.float(), ref, rtol=2 ** -10, atol=1e-6)
    B, T, H = 2, 197, 12
    qkv = torch.randn(B * T, 3 * H * 64, device="cuda", generator=g).half()
    o = eng.op_attention(qkv, B, T, H)
    q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    r = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(B * T, H * 64)
    assert torch.allclose(o.float(), r, rtol=2 ** -8, atol=3e-3), (o.float() - r).abs().max()


def test_classify_host_stream_matches_blocking_call(engine):
    """the pipelined host-buffer API returns, batch by batch and in order, what the blocking call returns"""
    from importlib import import_module
    clipc = import_module("ai-interior-image-classifier_b200.clip_compat")
    vis = clipc.build_visual("ViT-B/16", seed=0).cuda()
    eng = vis.sync_engine()
    g = torch.Generator().manual_seed(31)
    text = torch.nn.functional.normalize(torch.randn(60, 512, generator=g), dim=-1).cuda()
    eng.set_labels(text, [40, 20], [11, 0], topk=5, logit_scale=100.0)
    batches = [torch.randint(0, 256, (6, 224, 224, 3), dtype=torch.uint8, generator=g).pin_memory() for _ in range(5)]
    want = [eng.classify_host_u8(b) for b in batches]
    got = list(eng.classify_host_stream(iter(batches)))
    assert len(got) == len(want)
    for (tv, ti, ss), (wv, wi, ws) in zip(got, want):
        assert torch.equal(ti, wi) and torch.equal(tv, wv) and torch.equal(ss, ws)


def test_small_batch_cuda_graph_matches_direct_launches(engine):
    """batches <= Engine.graph_max_batch replay a CUDA graph of the whole path: same bits as the direct launches, also after the
    labels or a LoRA pair changed (the graphs are dropped then) and when the input buffer moves"""
    from importlib import import_module
    clipc = import_module("ai-interior-image-classifier_b200.clip_compat")
    lora = import_module("ai-interior-image-classifier_b200.lora")
    vis = clipc.build_visual("ViT-B/16", seed=0).cuda()
    eng = vis.sync_engine()
    g = torch.Generator().manual_seed(41)
    text = torch.nn.functional.normalize(torch.randn(60, 512, generator=g), dim=-1).cuda()
    eng.set_labels(text, [40, 20], [11, 0], topk=5, logit_scale=100.0)
    for B in (1, 5, 16):
        imgs = torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8, generator=g).cuda()
        want = eng.classify_same_size(imgs, use_graph=False)
        for _ in range(2):                                   # capture, then replay
            got = eng.classify_same_size(imgs.clone())       # default: graph for B <= 16
            assert torch.equal(got.logits, want.logits) and torch.equal(got.topk_idx, want.topk_idx)
            assert torch.equal(got.embedding, want.embedding) and torch.equal(got.split_sum, want.split_sum)
    assert set(eng._graphs) == {1, 5, 16}
    imgs = torch.randint(0, 256, (5, 224, 224, 3), dtype=torch.uint8, generator=g).cuda()
    # new labels -> graphs dropped, results follow
    text2 = torch.nn.functional.normalize(torch.randn(60, 512, generator=g), dim=-1).cuda()
    eng.set_labels(text2, [40, 20], [11, 0], topk=5, logit_scale=100.0)
    assert not eng._graphs
    assert torch.equal(eng.classify_same_size(imgs).logits, eng.classify_same_size(imgs, use_graph=False).logits)
    # a LoRA pair appears -> graphs dropped, results follow
    blk = vis.transformer.resblocks[3]
    blk.mlp.c_fc = lora.LoRALinear(blk.mlp.c_fc, rank=4, alpha=8)
    blk.mlp.c_fc.lora.lora_B.data.normal_(0, 0.02)
    eng = vis.sync_engine()
    a = eng.classify_same_size(imgs)
    b = eng.classify_same_size(imgs, use_graph=False)
    assert torch.equal(a.logits, b.logits) and not torch.equal(a.logits, eng.classify_same_size(imgs, use_graph=False).logits * 0)
