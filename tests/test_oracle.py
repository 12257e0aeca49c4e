"""Pins the ORACLE (oracle/): the CPU restatement must reproduce (i) the golden outputs that the reference's own code
produced in the build container, (ii) an independent implementation of the same published architecture (HF
transformers CLIP), and (iii) the known answers derivable from the repo's artefacts.  CPU only."""
import copy
import json
import os

import numpy as np
import pytest
import torch

from _common import GOLDEN, golden_json, golden_npz, oracle_model

torch.set_num_threads(max(1, os.cpu_count() or 1))
MEAN = torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(1, 3, 1, 1)
STD = torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(1, 3, 1, 1)


def _x(n, start=0):
    crops = torch.from_numpy(golden_npz("crops_u8.npz")["crops"][start:start + n])
    return (crops.permute(0, 3, 1, 2).float() / 255 - MEAN) / STD


def test_seeded_weights_are_the_golden_ones():
    from oracle import ref_semantics as RS
    meta = golden_json("meta.json")
    got = RS.weights_checksum(oracle_model())
    for k, v in meta["weights_checksum"].items():
        assert abs(got[k] - v) <= 1e-9 * abs(v), k
    assert meta["loader_kat"] == {"wrapped": 72, "loaded": 48, "missing": 96, "visual_lora_B_all_zero": 1}


def test_oracle_reproduces_reference_outputs():
    """embeddings + logits + detector + top-5 of the first images == what /root/reference/main.py produced"""
    from oracle import ref_semantics as RS
    m, ref, lab = oracle_model(), golden_npz("ref_shipped.npz"), golden_json("labels.json")
    text = torch.from_numpy(golden_npz("text_features.npz")["text"])
    with torch.no_grad():
        emb = m.encode_image(_x(4))
    assert torch.allclose(emb, torch.from_numpy(ref["emb"][:4]), rtol=1e-4, atol=1e-4)
    f = emb / emb.norm(dim=-1, keepdim=True)
    logits = 100.0 * f @ text.T
    assert np.allclose(logits.numpy(), ref["logits"][:4], atol=2e-3)
    top5 = json.loads(str(ref["top5"]))
    for i in range(4):
        d = RS.detector_decision(logits[i, :40], lab["detector"], 0.3)
        assert d[0] == bool(ref["det_is"][i]) and d[2] == str(ref["det_cat"][i]) and abs(d[1] - ref["det_conf"][i]) < 1e-4
        got = RS.group_topk(logits[i, 40:], lab["groups"])
        for g in lab["group_order"]:
            assert [l for l, _ in got[g]] == [l for l, _ in top5[i][g]]


def test_out_proj_lora_is_dead_and_mlp_lora_is_live():
    """SURVEY F3/F4 on the restated reference LoRA code: 72 wraps; out_proj LoRA never reaches the output."""
    from oracle import ref_semantics as RS
    m = copy.deepcopy(oracle_model())
    x = _x(2)
    with torch.no_grad():
        base = m.encode_image(x)
        assert len(RS.replace_linears_with_lora(m, rank=4, alpha=8)) == 72
        assert torch.equal(m.encode_image(x), base)                      # lora_B == 0 -> exact no-op
        g = torch.Generator().manual_seed(5)
        for n, p in m.named_parameters():
            if n.startswith("visual.") and n.endswith("attn.out_proj.lora.lora_B"):
                p.data = torch.randn(p.shape, generator=g)
        assert torch.equal(m.encode_image(x), base)                      # F4
        for n, p in m.named_parameters():
            if n.startswith("visual.") and n.endswith("mlp.c_fc.lora.lora_B"):
                p.data = torch.randn(p.shape, generator=g) * 0.02
        assert not torch.allclose(m.encode_image(x), base, atol=1e-3)


def test_vision_lora_golden():
    from oracle import ref_semantics as RS
    m = copy.deepcopy(oracle_model())
    RS.replace_linears_with_lora(m, rank=4, alpha=8)
    RS.seed_vision_lora(m, seed=1234)
    with torch.no_grad():
        emb = m.encode_image(_x(2, start=7))
    assert torch.allclose(emb, torch.from_numpy(golden_npz("ref_visionlora.npz")["emb"][7:9]), rtol=1e-4, atol=1e-4)


def test_against_independent_implementation_hf():
    """Same tensors in transformers' CLIP (quick_gelu, eps 1e-5, patch 16): fp32 agreement of both towers."""
    tr = pytest.importorskip("transformers")
    from oracle import clip_ref
    m = oracle_model()
    sd = m.state_dict()
    vc = tr.CLIPVisionConfig(hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                             image_size=224, patch_size=16, hidden_act="quick_gelu", layer_norm_eps=1e-5, projection_dim=512)
    tc = tr.CLIPTextConfig(vocab_size=49408, hidden_size=512, intermediate_size=2048, num_hidden_layers=12,
                           num_attention_heads=8, max_position_embeddings=77, hidden_act="quick_gelu", layer_norm_eps=1e-5,
                           projection_dim=512, eos_token_id=49407, bos_token_id=49406, pad_token_id=0)
    hv = tr.CLIPVisionModelWithProjection(vc).eval()
    ht = tr.CLIPTextModelWithProjection(tc).eval()

    def blocks(prefix_src, dst_layers, width):
        for i, lyr in enumerate(dst_layers):
            p = f"{prefix_src}.resblocks.{i}."
            w, b = sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"]
            for j, name in enumerate(("q_proj", "k_proj", "v_proj")):
                getattr(lyr.self_attn, name).weight.data = w[j * width:(j + 1) * width].clone()
                getattr(lyr.self_attn, name).bias.data = b[j * width:(j + 1) * width].clone()
            lyr.self_attn.out_proj.weight.data = sd[p + "attn.out_proj.weight"].clone()
            lyr.self_attn.out_proj.bias.data = sd[p + "attn.out_proj.bias"].clone()
            lyr.layer_norm1.weight.data, lyr.layer_norm1.bias.data = sd[p + "ln_1.weight"].clone(), sd[p + "ln_1.bias"].clone()
            lyr.layer_norm2.weight.data, lyr.layer_norm2.bias.data = sd[p + "ln_2.weight"].clone(), sd[p + "ln_2.bias"].clone()
            lyr.mlp.fc1.weight.data, lyr.mlp.fc1.bias.data = sd[p + "mlp.c_fc.weight"].clone(), sd[p + "mlp.c_fc.bias"].clone()
            lyr.mlp.fc2.weight.data, lyr.mlp.fc2.bias.data = sd[p + "mlp.c_proj.weight"].clone(), sd[p + "mlp.c_proj.bias"].clone()

    v = hv.vision_model
    v.embeddings.patch_embedding.weight.data = sd["visual.conv1.weight"].clone()
    v.embeddings.class_embedding.data = sd["visual.class_embedding"].clone()
    v.embeddings.position_embedding.weight.data = sd["visual.positional_embedding"].clone()
    v.pre_layrnorm.weight.data, v.pre_layrnorm.bias.data = sd["visual.ln_pre.weight"].clone(), sd["visual.ln_pre.bias"].clone()
    v.post_layernorm.weight.data, v.post_layernorm.bias.data = sd["visual.ln_post.weight"].clone(), sd["visual.ln_post.bias"].clone()
    hv.visual_projection.weight.data = sd["visual.proj"].t().clone()
    blocks("visual.transformer", v.encoder.layers, 768)
    t = ht.text_model
    t.embeddings.token_embedding.weight.data = sd["token_embedding.weight"].clone()
    t.embeddings.position_embedding.weight.data = sd["positional_embedding"].clone()
    t.final_layer_norm.weight.data, t.final_layer_norm.bias.data = sd["ln_final.weight"].clone(), sd["ln_final.bias"].clone()
    ht.text_projection.weight.data = sd["text_projection"].t().clone()
    blocks("transformer", t.encoder.layers, 512)
    x = _x(2)
    tok = clip_ref.tokenize(["wnętrze z drewno", "salon", "interior of a room"])
    with torch.no_grad():
        a, b = m.encode_image(x), hv(pixel_values=x).image_embeds
        c, d = m.encode_text(tok), ht(input_ids=tok).text_embeds
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-4), (a - b).abs().max()
    assert torch.allclose(c, d, rtol=1e-4, atol=1e-4), (c - d).abs().max()


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_restatement_equals_reference_code():
    """ref_semantics.{LoRALinear, replace_linears_with_lora, load_lora_state} vs /root/reference/main.py itself."""
    import sys
    from oracle import clip_ref, ref_semantics as RS
    clip_ref.install_clip_stub()
    sys.path.insert(0, "/root/reference")
    import main as ref_main
    a, b = copy.deepcopy(oracle_model()), copy.deepcopy(oracle_model())
    na = ref_main.replace_linears_with_lora(a, rank=4, alpha=8)
    nb = RS.replace_linears_with_lora(b, rank=4, alpha=8)
    assert na == nb and [n for n, _ in a.named_parameters()] == [n for n, _ in b.named_parameters()]
    ck = torch.load("/root/reference/lora_models/comprehensive_lora.pth", map_location="cpu")
    la = ref_main.load_lora_weights_to_model(a, "/root/reference/lora_models/comprehensive_lora.pth")
    lb = RS.load_lora_state(b, ck)
    assert la[0] == lb[0] == 48 and la[1] == lb[1] and len(la[1]) == 96
    tok = clip_ref.tokenize(["wnętrze z cegła"])
    with torch.no_grad():
        for n in ("transformer.resblocks.0.mlp.c_fc.lora.lora_A", "transformer.resblocks.11.mlp.c_proj.lora.lora_B"):
            assert torch.equal(dict(a.named_parameters())[n], dict(b.named_parameters())[n])
        assert torch.equal(a.encode_text(tok), b.encode_text(tok))


# ---------------------------------------------------------------------------------------------- JPEG ingest (row N2)
def _jpeg_fixtures():
    import json
    import numpy as np
    d = os.path.join(GOLDEN, "jpeg")
    meta = json.load(open(os.path.join(d, "meta.json")))["files"]
    exp = np.load(os.path.join(d, "expected.npz"))
    return d, meta, exp


def test_jpeg_oracle_matches_pillow_on_fixtures():
    """oracle/jpeg_ref.py == the committed Pillow outputs == Pillow on this box, bit for bit, on every fixture of the envelope;
    files outside it raise Unsupported"""
    import io
    import numpy as np
    from PIL import Image
    from oracle import jpeg_ref as J
    d, meta, exp = _jpeg_fixtures()
    seen = 0
    for name, m in meta.items():
        data = open(os.path.join(d, name + ".jpg"), "rb").read()
        live = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
        assert np.array_equal(live, exp[name]), name          # the committed vectors are what Pillow produces here too
        if m["in_envelope"]:
            got = J.decode_rgb(data)
            assert got.shape == exp[name].shape and np.array_equal(got, exp[name]), name
            seen += 1
        else:
            with pytest.raises(J.Unsupported):
                J.decode_rgb(data)
    assert seen >= 8


def test_jpeg_oracle_matches_pillow_on_fresh_encodes():
    """files Pillow encodes on the fly: every sampling layout, narrow / odd sizes (the box-replication rule for components of
    width <= 2), restart intervals, optimised Huffman tables, extreme qualities"""
    import io
    import numpy as np
    from PIL import Image
    from oracle import jpeg_ref as J
    rng = np.random.default_rng(7)
    n = 0
    for (h, w) in ((1, 1), (2, 5), (8, 8), (9, 17), (16, 3), (31, 33), (40, 4)):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        for sub in (0, 1, 2):
            for kw in (dict(quality=92), dict(quality=15, optimize=True), dict(quality=100, restart_marker_blocks=2)):
                buf = io.BytesIO()
                Image.fromarray(img).save(buf, "JPEG", subsampling=sub, **kw)
                data = buf.getvalue()
                ref = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
                assert np.array_equal(J.decode_rgb(data), ref), (h, w, sub, kw)
                n += 1
    assert n == 63


def test_jpeg_oracle_pinned_on_reference_dataset():
    """oracle/pin_jpeg.py ran over the reference's own 151 files in the build container: 140 baseline files bit-exact with Pillow,
    11 progressive files outside the envelope, no mismatch"""
    import json
    pin = json.load(open(os.path.join(GOLDEN, "jpeg_pin.json")))["summary"]
    assert pin["files"] == 151 and pin["mismatch"] == 0 and pin["bit_exact"] == 140 and pin["outside_envelope"] == 11


def test_jpeg_plan_parses_headers_without_a_gpu():
    """the C-ABI plan object is host-only: sizes / status of every fixture agree with Pillow, and the layout sizes are sane"""
    import numpy as np
    import torch
    from importlib import import_module
    jp = import_module("ai-interior-image-classifier_b200.jpeg")
    d, meta, exp = _jpeg_fixtures()
    names = sorted(meta)
    files = [open(os.path.join(d, n + ".jpg"), "rb").read() for n in names] + [b"", b"\xff\xd8\xff", b"not a jpeg at all"]
    offsets = np.zeros(len(files) + 1, dtype=np.int64)
    np.cumsum([len(f) for f in files], out=offsets[1:])
    blob = torch.frombuffer(bytearray(b"".join(files) + bytes(64)), dtype=torch.uint8)
    plan = jp.JpegPlan(blob, offsets)
    for i, n in enumerate(names):
        if meta[n]["in_envelope"]:
            assert plan.status[i] == jp.JPEG_OK and plan.sizes[i] == exp[n].shape[:2], n
        else:
            assert plan.status[i] == jp.JPEG_UNSUPPORTED and plan.sizes[i] == (0, 0) and plan.reasons[i], n
    assert plan.status[-3:] == [jp.JPEG_CORRUPT] * 3
    assert plan.scratch_bytes > plan.staging_bytes > 8 * 640 + 2448     # 8 image descriptors + at least one Huffman table
    plan.close()


CRAFTED = [dict(width=64, height=48), dict(width=97, height=61, sampling=(2, 1)), dict(width=40, height=40, sampling=(1, 1)),
           dict(width=80, height=72, restart=3, fill_before_rst=2), dict(width=80, height=72, restart=1), dict(width=50, height=30, dqt16=True),
           dict(width=33, height=17, comp_ids=(0, 1, 2)), dict(width=48, height=48, comp_ids=(82, 71, 66)), dict(width=48, height=48, adobe=1),
           dict(width=48, height=48, jfif=False, adobe=1), dict(width=64, height=64, table_ids=((1, 1), (0, 0), (0, 0))),
           dict(width=64, height=64, redefine=True), dict(width=72, height=40, gray=True), dict(width=64, height=48, deep=False)]


def test_jpeg_oracle_matches_pillow_on_crafted_streams():
    """files written by tests/jpeg_craft.py with the features third-party encoders use and Pillow's encoder does not (16-bit Huffman
    codes, restart intervals + fill bytes, 16-bit DQT, redefined tables, unusual ids, Adobe / JFIF combinations, ZRL, blocks filled
    to coefficient 63): Pillow decodes them, the oracle agrees bit for bit; RGB-coded files are outside the envelope"""
    import io
    import numpy as np
    from PIL import Image
    import jpeg_craft as C
    from oracle import jpeg_ref as J
    rng = np.random.default_rng(0)
    for kw in CRAFTED:
        data = C.random_case(rng, **kw)
        ref = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
        assert np.array_equal(J.decode_rgb(data), ref), kw
    for kw in (dict(jfif=False, adobe=0), dict(jfif=False, comp_ids=(82, 71, 66))):
        with pytest.raises(J.Unsupported):
            J.decode_rgb(C.random_case(rng, width=32, height=32, **kw))
