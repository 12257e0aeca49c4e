"""Host-side driver of the CUDA engine: owns an `iic_handle`, keeps the device copies of the weights in the
layouts include/iic.h documents, and exposes the hot path as tensor-in / tensor-out calls.

PyTorch is used for device memory, streams and the one-off weight layout conversion only; every FLOP of the
image path runs in the kernels behind the C ABI.  There is no CPU path: constructing an engine without a CUDA
device, or without the built extension, raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import threading
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L


@dataclass(frozen=True)
class VisionArch:
    image_size: int = 224
    patch_size: int = 16
    width: int = 768
    layers: int = 12
    heads: int = 12
    embed_dim: int = 512
    activation: int = L.ACT_QUICK_GELU
    # sequence (text-tower) engines: seq_tokens = context length (77), causal attention; no patch embedding / ln_pre
    seq_tokens: int = 0
    causal: bool = False

    @property
    def mlp_dim(self) -> int:
        return 4 * self.width

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size

    @property
    def tokens(self) -> int:
        return self.seq_tokens if self.seq_tokens > 0 else self.grid * self.grid + 1


VIT_B_16 = VisionArch()
VIT_L_14_336 = VisionArch(image_size=336, patch_size=14, width=1024, layers=24, heads=16, embed_dim=768)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


@dataclass
class HeadResult:
    """Device tensors produced by the fused head kernel."""
    embedding: Optional[torch.Tensor]  # [B, E] fp32, un-normalised (== encode_image output)
    logits: torch.Tensor               # [B, L] fp32 = logit_scale * cos
    probs: torch.Tensor                # [B, L] fp32, softmax inside each label group
    topk_val: torch.Tensor             # [B, G, k] fp32
    topk_idx: torch.Tensor             # [B, G, k] int32, index inside the group (-1 = padding)
    split_sum: torch.Tensor            # [B, G] fp32


class Engine:
    """One engine = one GPU.  Calls are serialised by an internal lock (the reference calls its detector from a
    4-thread pool, /root/reference/main.py:345-346)."""

    def __init__(self, arch: VisionArch = VIT_B_16, device: "torch.device | str | int" = "cuda", gemm_ctas: int = 0,
                 operand_dtype: "torch.dtype | str | None" = None):
        self.lib = L.load()
        if not torch.cuda.is_available():
            raise RuntimeError("iic-b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if self.device.type != "cuda":
            raise RuntimeError(f"iic-b200 engine cannot run on device {self.device}: CUDA only")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.arch = arch
        # None -> IIC_OPERAND_DTYPE or _lib.DEFAULT_OPERAND_DTYPE (fp16): ONE default for engine, bench, smoke and tests
        operand_dtype = torch.float16 if L.operand_dtype_name(operand_dtype) == "f16" else torch.bfloat16
        self.op_dtype = operand_dtype
        cfg = L.IicConfig(arch.image_size, arch.patch_size, arch.width, arch.layers, arch.heads, arch.mlp_dim,
                          arch.embed_dim, arch.activation, self.device.index, gemm_ctas,
                          L.DTYPE_F16 if operand_dtype == torch.float16 else L.DTYPE_BF16, int(arch.seq_tokens),
                          1 if arch.causal else 0)
        h = C.c_void_p()
        rc = self.lib.iic_create(C.byref(h), C.byref(cfg))
        if rc != L.IIC_OK:
            msg = self.lib.iic_last_error(None)
            raise RuntimeError(f"iic_create failed (code {rc}): {msg.decode() if msg else '?'}")
        self.h = h
        dims = L.IicDims()
        L.check(self.h, self.lib.iic_get_dims(self.h, C.byref(dims)), "iic_get_dims")
        self.dims = dims
        self._lock = threading.RLock()
        self._weights: Dict[str, torch.Tensor] = {}    # keeps borrowed buffers alive
        self._lora: Dict[Tuple[int, int], Tuple[torch.Tensor, torch.Tensor]] = {}
        self._labels: Optional[Tuple[torch.Tensor, List[int], int]] = None
        self._workspace: Optional[torch.Tensor] = None
        self._patches: Optional[torch.Tensor] = None
        # CUDA graphs of the whole path for small batches (latency path): keyed by batch size, dropped whenever a pointer the
        # captured kernels read could have changed (weights, LoRA operands, labels)
        self._graphs: Dict[int, tuple] = {}
        self._profiling = False
        self.graph_max_batch = int(os.environ.get("IIC_GRAPH_MAX_BATCH", "16"))

    def __deepcopy__(self, memo):
        """An engine is bound to one device handle: a deep-copied model gets None here and builds its own engine on first use
        (clip_compat.VisionTransformer.engine), e.g. the per-device replicas of analyzer.analyze_images_batch(devices=...)."""
        return None

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.iic_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def load_weight(self, name: str, tensor: torch.Tensor) -> None:
        """`tensor` must already have the dtype/shape include/iic.h lists for `name`."""
        t = tensor.detach().to(self.device).contiguous()
        dt = {torch.float32: L.DTYPE_F32, torch.bfloat16: L.DTYPE_BF16, torch.float16: L.DTYPE_F16}[t.dtype]
        shape = (C.c_int64 * t.dim())(*t.shape)
        with self._lock:
            L.check(self.h, self.lib.iic_load_weight(self.h, name.encode(), t.data_ptr(), dt, t.dim(), shape),
                    f"iic_load_weight({name})")
            self._weights[name] = t
            self._graphs.clear()

    def load_visual_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        """`sd`: OpenAI-CLIP `visual.*` tensors keyed WITHOUT the `visual.` prefix (fp32, any device).
        Linear / conv weights are converted to the 16-bit operand dtype here, everything else stays fp32."""
        a = self.arch
        f32 = lambda t: t.detach().to(self.device, torch.float32).contiguous()
        bf16 = lambda t: t.detach().to(self.device, torch.float32).to(self.op_dtype).contiguous()
        conv = sd["conv1.weight"].detach().to(self.device, torch.float32).reshape(a.width, -1)
        if conv.shape[1] != self.dims.patch_kpad:
            conv = torch.nn.functional.pad(conv, (0, self.dims.patch_kpad - conv.shape[1]))
        self.load_weight("conv1.weight", conv.to(self.op_dtype))
        for n in ("class_embedding", "positional_embedding", "ln_pre.weight", "ln_pre.bias", "ln_post.weight",
                  "ln_post.bias", "proj"):
            self.load_weight(n, f32(sd[n]))
        for i in range(a.layers):
            p = f"transformer.resblocks.{i}."
            for n in ("ln_1.weight", "ln_1.bias", "ln_2.weight", "ln_2.bias", "attn.in_proj_bias", "attn.out_proj.bias",
                      "mlp.c_fc.bias", "mlp.c_proj.bias"):
                self.load_weight(p + n, f32(sd[p + n]))
            for n in ("attn.in_proj_weight", "attn.out_proj.weight", "mlp.c_fc.weight", "mlp.c_proj.weight"):
                self.load_weight(p + n, bf16(sd[p + n]))

    def load_text_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        """Text tower on a sequence engine (arch.seq_tokens > 0).  `sd`: `transformer.resblocks.*`, `ln_final.*` and
        `text_projection` of OpenAI CLIP (fp32, any device); ln_final / text_projection occupy the engine's ln_post / proj slots."""
        a = self.arch
        f32 = lambda t: t.detach().to(self.device, torch.float32).contiguous()
        w16 = lambda t: t.detach().to(self.device, torch.float32).to(self.op_dtype).contiguous()
        self.load_weight("ln_post.weight", f32(sd["ln_final.weight"]))
        self.load_weight("ln_post.bias", f32(sd["ln_final.bias"]))
        self.load_weight("proj", f32(sd["text_projection"]))
        for i in range(a.layers):
            p = f"transformer.resblocks.{i}."
            for n in ("ln_1.weight", "ln_1.bias", "ln_2.weight", "ln_2.bias", "attn.in_proj_bias", "attn.out_proj.bias",
                      "mlp.c_fc.bias", "mlp.c_proj.bias"):
                self.load_weight(p + n, f32(sd[p + n]))
            for n in ("attn.in_proj_weight", "attn.out_proj.weight", "mlp.c_fc.weight", "mlp.c_proj.weight"):
                self.load_weight(p + n, w16(sd[p + n]))

    def set_lora(self, layer: int, which: int, lora_a: Optional[torch.Tensor], lora_b: Optional[torch.Tensor],
                 scaling: float = 1.0) -> None:
        """lora_a [in, r], lora_b [r, out] exactly as the reference stores them (main.py:26-27); `scaling` is
        alpha / rank (main.py:28).  None clears the slot."""
        with self._lock:
            self._graphs.clear()
            if lora_a is None or lora_b is None:
                L.check(self.h, self.lib.iic_set_lora(self.h, layer, which, None, None, 0), "iic_set_lora")
                self._lora.pop((layer, which), None)
                return
            r = lora_a.shape[1]
            d, mlp = self.arch.width, self.arch.mlp_dim
            want_in, want_out = {L.LORA_IN_PROJ: (d, 3 * d), L.LORA_OUT_PROJ: (d, d), L.LORA_C_FC: (d, mlp),
                                 L.LORA_C_PROJ: (mlp, d)}[which]
            if tuple(lora_a.shape) != (want_in, r) or tuple(lora_b.shape) != (r, want_out):
                raise ValueError(
                    f"LoRA pair for layer {layer} slot {which} has shapes {tuple(lora_a.shape)} / {tuple(lora_b.shape)}, "
                    f"expected ({want_in}, r) / (r, {want_out}) - a text-tower checkpoint suffix-matched onto the vision "
                    f"tower? (SURVEY appendix B hazard)")
            r4 = (r + 3) // 4 * 4
            pad = self.dims.lora_pad
            a = torch.zeros(lora_a.shape[0], r4, device=self.device, dtype=torch.float32)
            a[:, :r] = lora_a.detach().to(self.device, torch.float32) * float(scaling)
            bt = torch.zeros(lora_b.shape[1], pad, device=self.device, dtype=self.op_dtype)
            bt[:, :r] = lora_b.detach().to(self.device, torch.float32).t().to(self.op_dtype)
            L.check(self.h, self.lib.iic_set_lora(self.h, layer, which, a.data_ptr(), bt.data_ptr(), r), "iic_set_lora")
            self._lora[(layer, which)] = (a, bt)
            if r > 4:
                # ranks above 4: the down-projection x . A runs on the tcgen05 GEMM and wants (s A)^T in the operand format
                at16 = torch.zeros(pad, lora_a.shape[0], device=self.device, dtype=self.op_dtype)
                at16[:r] = a[:, :r].t().to(self.op_dtype)
                L.check(self.h, self.lib.iic_set_lora_operands16(self.h, layer, which, at16.data_ptr(), None),
                        "iic_set_lora_operands16")
                self._lora[(layer, which)] = (a, bt, at16)

    def set_labels(self, text_features: torch.Tensor, group_sizes: Sequence[int],
                   group_split: Optional[Sequence[int]] = None, topk: int = 5, logit_scale: float = 100.0) -> None:
        """text_features [L, E]: L2-normalised label embeddings, groups concatenated in order."""
        t = text_features.detach().to(self.device, torch.float32).contiguous()
        offs = [0]
        for s in group_sizes:
            offs.append(offs[-1] + int(s))
        if offs[-1] != t.shape[0]:
            raise ValueError("group sizes do not add up to the number of label rows")
        G = len(group_sizes)
        c_off = (C.c_int * (G + 1))(*offs)
        c_split = (C.c_int * G)(*[int(x) for x in group_split]) if group_split is not None else None
        with self._lock:
            L.check(self.h, self.lib.iic_set_labels(self.h, t.data_ptr(), t.shape[0], c_off, c_split, G, int(topk),
                                                    float(logit_scale)), "iic_set_labels")
            self._labels = (t, offs, int(topk))
            self._graphs.clear()

    # ------------------------------------------------------------------ buffers
    def _ws(self, B: int) -> torch.Tensor:
        need = int(self.lib.iic_workspace_bytes(self.h, B))
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = None
            self._workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._workspace

    def patch_buffer(self, B: int) -> torch.Tensor:
        rows = B * self.arch.grid * self.arch.grid
        if self._patches is None or self._patches.shape[0] < rows:
            self._patches = None
            self._patches = torch.zeros(rows, self.dims.patch_kpad, dtype=self.op_dtype, device=self.device)
        return self._patches[:rows]

    # ------------------------------------------------------------------ preprocessing
    def preprocess_same_size(self, images_u8: torch.Tensor, out: Optional[torch.Tensor] = None,
                             layout: int = L.OUT_PATCHES_BF16) -> torch.Tensor:
        """uint8 [B, R, R, 3] on the device -> patch matrix (default) or CHW tensor."""
        a = self.arch
        assert images_u8.dtype == torch.uint8 and images_u8.is_cuda and images_u8.is_contiguous()
        B = images_u8.shape[0]
        assert tuple(images_u8.shape[1:]) == (a.image_size, a.image_size, 3), images_u8.shape
        if out is None:
            out = self._alloc_pre_out(B, layout)
        with self._lock, torch.cuda.device(self.device):
            L.check(self.h, self.lib.iic_preprocess_same_size(self.h, images_u8.data_ptr(), B, out.data_ptr(), layout,
                                                              _stream_ptr(self.device)), "iic_preprocess_same_size")
        return out

    def _alloc_pre_out(self, B: int, layout: int) -> torch.Tensor:
        a = self.arch
        if layout == L.OUT_PATCHES_BF16:
            return self.patch_buffer(B)
        dt = torch.float32 if layout == L.OUT_CHW_F32 else self.op_dtype
        return torch.empty(B, 3, a.image_size, a.image_size, dtype=dt, device=self.device)

    def preprocess(self, images_u8: Sequence[torch.Tensor], out: Optional[torch.Tensor] = None,
                   layout: int = L.OUT_PATCHES_BF16) -> torch.Tensor:
        """List of uint8 [H, W, 3] device tensors of any size -> PIL-compatible resize + crop + normalise."""
        B = len(images_u8)
        for t in images_u8:
            assert t.dtype == torch.uint8 and t.is_cuda and t.is_contiguous() and t.dim() == 3 and t.shape[2] == 3
        ptrs = (C.c_void_p * B)(*[t.data_ptr() for t in images_u8])
        hw = (C.c_int * (2 * B))(*[v for t in images_u8 for v in (t.shape[0], t.shape[1])])
        if out is None:
            out = self._alloc_pre_out(B, layout)
        if B == 0:
            return out
        with self._lock, torch.cuda.device(self.device):
            L.check(self.h, self.lib.iic_preprocess(self.h, ptrs, hw, B, out.data_ptr(), layout,
                                                    _stream_ptr(self.device)), "iic_preprocess")
        return out

    def patchify(self, chw: torch.Tensor) -> torch.Tensor:
        a = self.arch
        assert chw.is_cuda and chw.dim() == 4 and tuple(chw.shape[1:]) == (3, a.image_size, a.image_size), chw.shape
        chw = chw.contiguous()
        dt = {torch.float32: L.DTYPE_F32, torch.bfloat16: L.DTYPE_BF16, torch.float16: L.DTYPE_F16}[chw.dtype]
        out = self.patch_buffer(chw.shape[0])
        with self._lock, torch.cuda.device(self.device):
            L.check(self.h, self.lib.iic_patchify(self.h, chw.data_ptr(), dt, chw.shape[0], out.data_ptr(),
                                                  _stream_ptr(self.device)), "iic_patchify")
        return out

    # ------------------------------------------------------------------ encoder / head
    def encode_patches(self, patches: torch.Tensor, B: int) -> torch.Tensor:
        emb = torch.empty(B, self.arch.embed_dim, dtype=torch.float32, device=self.device)
        with self._lock, torch.cuda.device(self.device):
            ws = self._ws(B)
            L.check(self.h, self.lib.iic_encode(self.h, patches.data_ptr(), B, ws.data_ptr(), ws.numel(),
                                                emb.data_ptr(), _stream_ptr(self.device)), "iic_encode")
        return emb

    def encode_sequence(self, x: torch.Tensor, row_index: torch.Tensor) -> torch.Tensor:
        """x f32 [B, T, width] (token + positional embedding), row_index [B] (EOT position) -> [B, E] fp32, un-normalised:
        the engine behind model.encode_text (residual blocks with the causal mask, ln_final of the EOT row, text_projection)."""
        if self.arch.seq_tokens <= 0:
            raise RuntimeError("encode_sequence needs an engine built with VisionArch(seq_tokens=...)")
        B, T, d = x.shape
        if T != self.arch.seq_tokens or d != self.arch.width:
            raise ValueError(f"expected [B, {self.arch.seq_tokens}, {self.arch.width}], got {tuple(x.shape)}")
        x = x.detach().to(self.device, torch.float32).contiguous()
        idx = row_index.detach().to(self.device, torch.int32).contiguous()
        emb = torch.empty(B, self.arch.embed_dim, dtype=torch.float32, device=self.device)
        with self._lock, torch.cuda.device(self.device):
            ws = self._ws(B)
            L.check(self.h, self.lib.iic_encode_sequence(self.h, x.data_ptr(), idx.data_ptr(), B, ws.data_ptr(), ws.numel(),
                                                         emb.data_ptr(), _stream_ptr(self.device)), "iic_encode_sequence")
        return emb

    def encode_image(self, chw: torch.Tensor) -> torch.Tensor:
        """[B,3,R,R] float tensor -> [B,E] fp32 (un-normalised): the engine behind model.encode_image."""
        with self._lock:
            return self.encode_patches(self.patchify(chw), chw.shape[0])

    def _head_out(self, B: int, want_logits: bool = True):
        if self._labels is None:
            raise RuntimeError("set_labels() has not been called")
        t, offs, k = self._labels
        Lc, G = t.shape[0], len(offs) - 1
        dev = self.device
        logits = torch.empty(B, Lc, dtype=torch.float32, device=dev)
        probs = torch.empty(B, Lc, dtype=torch.float32, device=dev)
        tv = torch.empty(B, G, k, dtype=torch.float32, device=dev)
        ti = torch.empty(B, G, k, dtype=torch.int32, device=dev)
        ss = torch.empty(B, G, dtype=torch.float32, device=dev)
        out = L.IicHeadOut(logits.data_ptr(), probs.data_ptr(), tv.data_ptr(), ti.data_ptr(), ss.data_ptr())
        return out, logits, probs, tv, ti, ss

    def head(self, emb: torch.Tensor) -> HeadResult:
        emb = emb.detach().to(self.device, torch.float32).contiguous()
        B = emb.shape[0]
        with self._lock, torch.cuda.device(self.device):
            out, logits, probs, tv, ti, ss = self._head_out(B)
            L.check(self.h, self.lib.iic_head(self.h, emb.data_ptr(), B, C.byref(out), _stream_ptr(self.device)),
                    "iic_head")
        return HeadResult(emb, logits, probs, tv, ti, ss)

    def classify_patches(self, patches: torch.Tensor, B: int, want_embedding: bool = True) -> HeadResult:
        with self._lock, torch.cuda.device(self.device):
            out, logits, probs, tv, ti, ss = self._head_out(B)
            emb = torch.empty(B, self.arch.embed_dim, dtype=torch.float32, device=self.device) if want_embedding else None
            ws = self._ws(B)
            L.check(self.h, self.lib.iic_classify(self.h, patches.data_ptr(), B, ws.data_ptr(), ws.numel(), _ptr(emb),
                                                  C.byref(out), _stream_ptr(self.device)), "iic_classify")
        return HeadResult(emb, logits, probs, tv, ti, ss)

    def classify_same_size(self, images_u8: torch.Tensor, want_embedding: bool = True, use_graph: Optional[bool] = None) -> HeadResult:
        """uint8 [B,R,R,3] device tensor -> preprocess + encoder + head.  Small batches (<= graph_max_batch) replay a CUDA
        graph of the ~100 launches of the path (the single-image case of main.py is launch-bound otherwise)."""
        B = images_u8.shape[0]
        if use_graph is None:
            use_graph = 0 < B <= self.graph_max_batch and not self._profiling and not torch.cuda.is_current_stream_capturing()
        with self._lock:
            if use_graph:
                return self._classify_graphed(images_u8, want_embedding)
            return self.classify_patches(self.preprocess_same_size(images_u8), B, want_embedding)

    def _classify_graphed(self, images_u8: torch.Tensor, want_embedding: bool) -> HeadResult:
        B = images_u8.shape[0]
        assert images_u8.dtype == torch.uint8 and images_u8.is_cuda and images_u8.is_contiguous()
        assert tuple(images_u8.shape[1:]) == (self.arch.image_size, self.arch.image_size, 3), images_u8.shape
        entry = self._graphs.get(B)
        if entry is None:
            dev = self.device
            with torch.cuda.device(dev):
                # private, static buffers: nothing the captured kernels touch may move or be reused while the graph lives
                static_in = torch.empty_like(images_u8)
                patches = torch.zeros(B * self.arch.grid * self.arch.grid, self.dims.patch_kpad, dtype=self.op_dtype, device=dev)
                ws = torch.empty(int(self.lib.iic_workspace_bytes(self.h, B)), dtype=torch.uint8, device=dev)
                out, logits, probs, tv, ti, ss = self._head_out(B)
                emb = torch.empty(B, self.arch.embed_dim, dtype=torch.float32, device=dev)

                def run():
                    s = _stream_ptr(dev)
                    L.check(self.h, self.lib.iic_preprocess_same_size(self.h, static_in.data_ptr(), B, patches.data_ptr(),
                                                                      L.OUT_PATCHES_BF16, s), "iic_preprocess_same_size")
                    L.check(self.h, self.lib.iic_classify(self.h, patches.data_ptr(), B, ws.data_ptr(), ws.numel(), emb.data_ptr(),
                                                          C.byref(out), s), "iic_classify")

                static_in.copy_(images_u8)
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    run()                      # warm-up outside capture: one-off attribute / descriptor setup
                torch.cuda.current_stream(dev).wait_stream(side)
                torch.cuda.synchronize(dev)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    run()
            entry = (graph, static_in, (emb, logits, probs, tv, ti, ss), (patches, ws, out))
            self._graphs[B] = entry
        graph, static_in, (emb, logits, probs, tv, ti, ss), _keep = entry
        with torch.cuda.device(self.device):
            static_in.copy_(images_u8, non_blocking=True)
            graph.replay()
            # hand out copies: the static outputs are overwritten by the next replay
            return HeadResult(emb.clone() if want_embedding else None, logits.clone(), probs.clone(), tv.clone(), ti.clone(), ss.clone())

    def classify(self, images_u8: Sequence[torch.Tensor], want_embedding: bool = True) -> HeadResult:
        """list of uint8 [H,W,3] device tensors (any size) -> preprocess + encoder + head."""
        with self._lock:
            return self.classify_patches(self.preprocess(images_u8), len(images_u8), want_embedding)

    # ------------------------------------------------------------------ training (LoRA-only gradients)
    def enable_training(self, sd: Dict[str, torch.Tensor]) -> None:
        """Upload the transposed projection weights the dX GEMMs need (`sd` as in load_visual_state_dict)."""
        for i in range(self.arch.layers):
            p = f"transformer.resblocks.{i}."
            for n in ("attn.in_proj_weight", "attn.out_proj.weight", "mlp.c_fc.weight", "mlp.c_proj.weight"):
                w = sd[p + n].detach().to(self.device, torch.float32).to(self.op_dtype)
                self.load_weight(p + n + "_t", w.t().contiguous())

    def set_lora_train(self, layer: int, which: int, lora_a: torch.Tensor, lora_b: torch.Tensor, scaling: float,
                       grad_a: torch.Tensor, grad_b: torch.Tensor) -> None:
        """Extra operands + gradient destinations for one LoRA slot (after set_lora).  grad_a [in, r] and grad_b [r, out]
        are fp32 device tensors (typically views of the parameters' .grad) that train_backward overwrites."""
        r = lora_a.shape[1]
        r4 = (r + 3) // 4 * 4
        a16 = torch.zeros(lora_a.shape[0], self.dims.lora_pad, device=self.device, dtype=self.op_dtype)
        a16[:, :r] = (lora_a.detach().to(self.device, torch.float32) * float(scaling)).to(self.op_dtype)
        bt32 = torch.zeros(lora_b.shape[1], r4, device=self.device, dtype=torch.float32)
        bt32[:, :r] = lora_b.detach().to(self.device, torch.float32).t()
        for g, shape in ((grad_a, tuple(lora_a.shape)), (grad_b, tuple(lora_b.shape))):
            assert g.is_cuda and g.dtype == torch.float32 and g.is_contiguous() and tuple(g.shape) == shape
        with self._lock:
            L.check(self.h, self.lib.iic_set_lora_train(self.h, layer, which, a16.data_ptr(), bt32.data_ptr(), float(scaling),
                                                        grad_a.data_ptr(), grad_b.data_ptr()), "iic_set_lora_train")
            self._lora[(layer, which, "train")] = (a16, bt32, grad_a, grad_b)
            # 16-bit lora_B [lora_pad, out] for the backward's dP = dY . B^T on the tcgen05 GEMM (all ranks); (s A)^T exists for r > 4
            at16 = self._lora[(layer, which)][2] if r > 4 else None
            b16 = torch.zeros(self.dims.lora_pad, lora_b.shape[1], device=self.device, dtype=self.op_dtype)
            b16[:r] = lora_b.detach().to(self.device, torch.float32).to(self.op_dtype)
            L.check(self.h, self.lib.iic_set_lora_operands16(self.h, layer, which, _ptr(at16), b16.data_ptr()),
                    "iic_set_lora_operands16")
            self._lora[(layer, which, "train16")] = (b16,)

    def set_lora_source(self, layer: int, which: int, lora_a: torch.Tensor, lora_b: torch.Tensor, scaling: float) -> None:
        """Register the fp32 parameters of a slot (after set_lora [+ set_lora_train]) for `refresh_lora`."""
        for t in (lora_a, lora_b):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise ValueError("set_lora_source needs contiguous fp32 CUDA parameters")
        with self._lock:
            L.check(self.h, self.lib.iic_set_lora_source(self.h, layer, which, lora_a.data_ptr(), lora_b.data_ptr(), float(scaling)),
                    "iic_set_lora_source")
            self._lora[(layer, which, "src")] = (lora_a, lora_b)

    def refresh_lora(self) -> None:
        """Rebuild every derived LoRA operand from the registered parameters (one small kernel per slot, no host sync)."""
        with self._lock, torch.cuda.device(self.device):
            L.check(self.h, self.lib.iic_refresh_lora(self.h, _stream_ptr(self.device)), "iic_refresh_lora")

    def _train_ws(self, B: int) -> torch.Tensor:
        need = int(self.lib.iic_train_workspace_bytes(self.h, B))
        ws = getattr(self, "_train_workspace", None)
        if ws is None or ws.numel() < need:
            self._train_workspace = None
            self._train_workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._train_workspace

    def train_forward(self, patches: torch.Tensor, B: int) -> torch.Tensor:
        """-> x_cls f32 [B, width]: class-token rows of the final residual stream (input of ln_post)."""
        x_cls = torch.empty(B, self.arch.width, dtype=torch.float32, device=self.device)
        with self._lock, torch.cuda.device(self.device):
            ws = self._train_ws(B)
            L.check(self.h, self.lib.iic_train_forward(self.h, patches.data_ptr(), B, ws.data_ptr(), ws.numel(),
                                                       x_cls.data_ptr(), _stream_ptr(self.device)), "iic_train_forward")
        return x_cls

    def train_forward_sequence(self, x: torch.Tensor, row_index: torch.Tensor) -> torch.Tensor:
        """Text tower (sequence engine): x f32 [B, T, width] = token + positional embedding, row_index [B] = EOT positions ->
        f32 [B, width]: those rows of the final residual stream (input of ln_final); activations are kept for train_backward."""
        if self.arch.seq_tokens <= 0:
            raise RuntimeError("train_forward_sequence needs an engine built with VisionArch(seq_tokens=...)")
        B, T, d = x.shape
        if T != self.arch.seq_tokens or d != self.arch.width:
            raise ValueError(f"expected [B, {self.arch.seq_tokens}, {self.arch.width}], got {tuple(x.shape)}")
        x = x.detach().to(self.device, torch.float32).contiguous()
        idx = row_index.detach().to(self.device, torch.int32).contiguous()
        out = torch.empty(B, d, dtype=torch.float32, device=self.device)
        with self._lock, torch.cuda.device(self.device):
            ws = self._train_ws(B)
            L.check(self.h, self.lib.iic_train_forward_sequence(self.h, x.data_ptr(), idx.data_ptr(), B, ws.data_ptr(), ws.numel(),
                                                                out.data_ptr(), _stream_ptr(self.device)),
                    "iic_train_forward_sequence")
        return out

    def train_backward(self, dx_cls: torch.Tensor, layer_done=None, loss_scale: "float | str" = "auto",
                       row_index: Optional[torch.Tensor] = None) -> None:
        """dx_cls f32 [B, width] (sequence engines: the gradient of the rows `row_index` that train_forward_sequence returned).
        `layer_done(layer)` (optional) is called right after block `layer`'s backward has been
        enqueued - its LoRA gradients are final once the stream reaches that point (hook for the gradient all-reduce).
        loss_scale: power of two applied to the activation gradients and divided out of the LoRA gradients ("auto": chosen
        so that max |dx_cls| maps to 2^8, which keeps fp16 operands clear of both underflow and overflow)."""
        dx_cls = dx_cls.detach().to(self.device, torch.float32).contiguous()
        B = dx_cls.shape[0]
        if loss_scale == "auto":
            # max |dx_cls| of the PREVIOUS step picks the scale (no pipeline stall): it is copied into a PINNED buffer and an
            # event marks the copy, which is synchronised before the value is read - back-to-back forward_backward calls
            # (gradient accumulation, benches) therefore never read a half-written number.  The 2^8 target leaves 2^8 of
            # headroom below fp16's maximum for step-to-step drift; `loss_scale_backoff` (<= 1, halved by the trainers when a
            # step produced non-finite gradients, recovered slowly) is applied on top.
            with torch.cuda.device(self.device):
                if getattr(self, "_amax_pin", None) is None:
                    self._amax_pin = torch.zeros(1, dtype=torch.float32, pin_memory=True)
                    self._amax_event = None
                if self._amax_event is not None:
                    self._amax_event.synchronize()
                    amax = float(self._amax_pin[0])
                else:
                    amax = float(dx_cls.abs().max())
                self._amax_pin.copy_(dx_cls.abs().max().reshape(1), non_blocking=True)
                self._amax_event = torch.cuda.Event()
                self._amax_event.record(torch.cuda.current_stream(self.device))
            loss_scale = 2.0 ** math.floor(math.log2(256.0 / amax)) if amax > 0 and math.isfinite(amax) else 1.0
            loss_scale *= float(getattr(self, "loss_scale_backoff", 1.0))
        loss_scale = float(loss_scale)
        if loss_scale != 1.0:
            dx_cls = dx_cls * loss_scale
        with self._lock, torch.cuda.device(self.device):
            L.check(self.h, self.lib.iic_train_set_loss_scale(self.h, loss_scale), "iic_train_set_loss_scale")
            ws = self._train_ws(B)
            s = _stream_ptr(self.device)
            if layer_done is None and row_index is None:
                L.check(self.h, self.lib.iic_train_backward(self.h, B, ws.data_ptr(), ws.numel(), dx_cls.data_ptr(), s),
                        "iic_train_backward")
                return
            if row_index is not None:
                idx = row_index.detach().to(self.device, torch.int32).contiguous()
                L.check(self.h, self.lib.iic_train_backward_begin_sequence(self.h, B, ws.data_ptr(), ws.numel(), dx_cls.data_ptr(),
                                                                           idx.data_ptr(), s), "iic_train_backward_begin_sequence")
            else:
                L.check(self.h, self.lib.iic_train_backward_begin(self.h, B, ws.data_ptr(), ws.numel(), dx_cls.data_ptr(), s),
                        "iic_train_backward_begin")
            for layer in range(self.arch.layers - 1, -1, -1):
                L.check(self.h, self.lib.iic_train_backward_layer(self.h, B, ws.data_ptr(), ws.numel(), layer, s),
                        "iic_train_backward_layer")
                if layer_done is not None:
                    layer_done(layer)

    def op_attention_bwd(self, qkv: torch.Tensor, d_out: torch.Tensor, B: int, T: int, heads: int):
        out = torch.empty(B * T, heads * 64, dtype=self.op_dtype, device=self.device)
        dqkv = torch.empty_like(qkv)
        lse = torch.empty(2 * B * heads * T, dtype=torch.float32, device=self.device)   # lse + the backward's D scratch
        with self._lock, torch.cuda.device(self.device):
            L.check(self.h, self.lib.iic_op_attention_bwd(self.h, qkv.data_ptr(), out.data_ptr(), d_out.data_ptr(), dqkv.data_ptr(),
                                                          lse.data_ptr(), B, T, heads, _stream_ptr(self.device)), "iic_op_attention_bwd")
        return out, dqkv, lse[:B * heads * T]

    def op_layernorm_bwd(self, dy: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, dx: torch.Tensor):
        rows, D = x.shape
        dx16 = torch.empty(rows, D, dtype=self.op_dtype, device=self.device)
        with self._lock, torch.cuda.device(self.device):
            L.check(self.h, self.lib.iic_op_layernorm_bwd(self.h, dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), dx.data_ptr(),
                                                          dx16.data_ptr(), rows, D, _stream_ptr(self.device)), "iic_op_layernorm_bwd")
        return dx16

    def op_act_bwd(self, dh: torch.Tensor, u: torch.Tensor, act: int = 1) -> None:
        with self._lock, torch.cuda.device(self.device):
            L.check(self.h, self.lib.iic_op_act_bwd(self.h, dh.data_ptr(), u.data_ptr(), dh.numel(), act,
                                                    _stream_ptr(self.device)), "iic_op_act_bwd")

    def op_lora_bwd(self, P: torch.Tensor, Y: torch.Tensor, Bm: torch.Tensor, rank: int, scale: float = 1.0):
        """(dB f32 [rank, N] = scale * P^T . Y,  dP 16-bit [M, 16] = Y . Bm^T) from one pass over Y."""
        M, N = Y.shape
        db = torch.zeros(rank, N, dtype=torch.float32, device=self.device)
        dp = torch.zeros(M, 16, dtype=self.op_dtype, device=self.device)
        scratch = torch.empty(int(self.lib.iic_op_lora_scratch_bytes(N, M)), dtype=torch.uint8, device=self.device)   # caller-owned
        with self._lock, torch.cuda.device(self.device):
            L.check(self.h, self.lib.iic_op_lora_bwd(self.h, P.data_ptr(), P.stride(0), Y.data_ptr(), N, M, Bm.data_ptr(), rank,
                                                     float(scale), db.data_ptr(), dp.data_ptr(), scratch.data_ptr(), scratch.numel(),
                                                     _stream_ptr(self.device)), "iic_op_lora_bwd")
        return db, dp

    def op_lora_outer(self, P: torch.Tensor, Y: torch.Tensor, rank: int, act: int = 0, scale: float = 1.0,
                      transpose: bool = False) -> torch.Tensor:
        M, N = Y.shape
        out = torch.zeros((N, rank) if transpose else (rank, N), dtype=torch.float32, device=self.device)
        scratch = torch.empty(int(self.lib.iic_op_lora_scratch_bytes(N, M)), dtype=torch.uint8, device=self.device)   # caller-owned
        with self._lock, torch.cuda.device(self.device):
            L.check(self.h, self.lib.iic_op_lora_outer(self.h, P.data_ptr(), P.stride(0), Y.data_ptr(), N, M, act, rank, float(scale),
                                                       1 if transpose else 0, out.data_ptr(), scratch.data_ptr(), scratch.numel(),
                                                       _stream_ptr(self.device)), "iic_op_lora_outer")
        return out

    # ------------------------------------------------------------------ measurement
    def profile(self, enable: bool = True) -> None:
        self._profiling = bool(enable)
        L.check(self.h, self.lib.iic_profile(self.h, 1 if enable else 0), "iic_profile")

    def profile_read(self) -> Dict[str, Dict[str, float]]:
        """{kernel class: {"ms": summed CUDA-event time, "launches": count}} since the last read."""
        n = len(L.KERNEL_CLASSES)
        ms, cnt = (C.c_double * n)(), (C.c_longlong * n)()
        L.check(self.h, self.lib.iic_profile_read(self.h, ms, cnt, n), "iic_profile_read")
        return {k: {"ms": float(ms[i]), "launches": int(cnt[i])} for i, k in enumerate(L.KERNEL_CLASSES)}

    def classify_host_u8(self, images_u8_pinned: torch.Tensor, staging: Optional[torch.Tensor] = None):
        """End-to-end call with HOST buffers: uint8 [B,R,R,3] (pinned) -> H2D copy -> preprocess + encoder + head ->
        top-k values / indices / detector sums copied back to the host.  Returns host tensors."""
        with self._lock, torch.cuda.device(self.device):
            dev = images_u8_pinned.to(self.device, non_blocking=True) if staging is None else staging.copy_(
                images_u8_pinned, non_blocking=True)
            r = self.classify_same_size(dev, want_embedding=False)
            out = (r.topk_val.cpu(), r.topk_idx.cpu(), r.split_sum.cpu())
        return out

    def classify_host_stream(self, batches):
        """Streaming form of `classify_host_u8` for many batches: `batches` yields pinned uint8 [B,R,R,3] host tensors, the
        generator yields (topk_val, topk_idx, split_sum) host tensors in the same order.  Every batch still pays its own H2D
        copy and D2H read, but the copy of batch i+1 runs on a side stream while batch i is being encoded, and the host only
        waits for batch i-1 while batch i is in flight (two staging buffers, two pinned result sets)."""
        dev = self.device
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            copy = torch.cuda.Stream(device=dev)
            staging = [None, None]
            outs = [None, None]
            copied = [torch.cuda.Event(), torch.cuda.Event()]
            done = [torch.cuda.Event(), torch.cuda.Event()]
            pending = None   # slot whose results have been enqueued but not yet handed out
            for i, host in enumerate(batches):
                k = i & 1
                if staging[k] is None or staging[k].shape != host.shape:
                    staging[k] = torch.empty(host.shape, dtype=torch.uint8, device=dev)
                with torch.cuda.stream(copy):
                    if i >= 2:
                        copy.wait_event(done[k])          # the encode that read this staging buffer two batches ago
                    staging[k].copy_(host, non_blocking=True)
                    copied[k].record(copy)
                main.wait_event(copied[k])
                with self._lock:
                    r = self.classify_same_size(staging[k], want_embedding=False)
                    if outs[k] is None or outs[k][0].shape != r.topk_val.shape:
                        outs[k] = (torch.empty(r.topk_val.shape, dtype=r.topk_val.dtype, pin_memory=True),
                                   torch.empty(r.topk_idx.shape, dtype=r.topk_idx.dtype, pin_memory=True),
                                   torch.empty(r.split_sum.shape, dtype=r.split_sum.dtype, pin_memory=True))
                    outs[k][0].copy_(r.topk_val, non_blocking=True)
                    outs[k][1].copy_(r.topk_idx, non_blocking=True)
                    outs[k][2].copy_(r.split_sum, non_blocking=True)
                    done[k].record(main)
                if pending is not None:
                    done[pending].synchronize()
                    yield tuple(t.clone() for t in outs[pending])
                pending = k
            if pending is not None:
                done[pending].synchronize()
                yield tuple(t.clone() for t in outs[pending])

    # ------------------------------------------------------------------ single operators (tests / profiling)
    def op_gemm(self, a: torch.Tensor, w: torch.Tensor, epilogue: int, bias: Optional[torch.Tensor] = None,
                residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                lora_p: Optional[torch.Tensor] = None, lora_bt: Optional[torch.Tensor] = None, r_pad: int = 0,
                group: int = 1, ctas: int = 0, out_rows: Optional[int] = None, down_a: Optional[torch.Tensor] = None,
                down_part: Optional[torch.Tensor] = None) -> torch.Tensor:
        M, K = a.shape
        N = w.shape[0]
        f32_out = epilogue in (L.EPI_BIAS_RES_F32, L.EPI_POS_F32)
        if out is None:
            out = torch.empty(out_rows or M, N, dtype=torch.float32 if f32_out else self.op_dtype, device=self.device)
        lora_ld = lora_p.stride(0) if lora_p is not None else 0
        with self._lock, torch.cuda.device(self.device):
            L.check(self.h, self.lib.iic_op_gemm(self.h, a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), M, N, K,
                                                 _ptr(lora_p), _ptr(lora_bt), r_pad, lora_ld, epilogue, _ptr(bias),
                                                 _ptr(residual), out.data_ptr(), out.stride(0), group, ctas, _ptr(down_a),
                                                 _ptr(down_part), _stream_ptr(self.device)), "iic_op_gemm")
        return out

    def op_gemm_act_dual(self, a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], act: int = 0,
                         lora_p: Optional[torch.Tensor] = None, lora_bt: Optional[torch.Tensor] = None, r_pad: int = 0,
                         ctas: int = 0):
        """(act(a . w^T + bias), a . w^T + bias), both in the operand dtype, from one launch (training forward of c_fc)."""
        M, K = a.shape
        N = w.shape[0]
        out = torch.empty(M, N, dtype=self.op_dtype, device=self.device)
        pre = torch.empty(M, N, dtype=self.op_dtype, device=self.device)
        lora_ld = lora_p.stride(0) if lora_p is not None else 0
        with self._lock, torch.cuda.device(self.device):
            L.check(self.h, self.lib.iic_op_gemm_act_dual(
                self.h, a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), M, N, K, _ptr(lora_p), _ptr(lora_bt), r_pad,
                lora_ld, _ptr(bias), out.data_ptr(), pre.data_ptr(), act, ctas, _stream_ptr(self.device)), "iic_op_gemm_act_dual")
        return out, pre

    def op_layernorm(self, x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, out_dtype=None,
                     lora_a_scaled: Optional[torch.Tensor] = None, p_ld: int = 16):
        rows, D = x.shape
        out_dtype = out_dtype or self.op_dtype
        out = torch.empty(rows, D, dtype=out_dtype, device=self.device)
        p = None
        r4 = 0
        if lora_a_scaled is not None:
            r4 = lora_a_scaled.shape[1]
            p = torch.zeros(rows, p_ld, dtype=self.op_dtype, device=self.device)
        with self._lock, torch.cuda.device(self.device):
            L.check(self.h, self.lib.iic_op_layernorm(
                self.h, x.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                out.data_ptr() if out_dtype != torch.float32 else None,
                out.data_ptr() if out_dtype == torch.float32 else None, rows, D, _ptr(lora_a_scaled), r4, _ptr(p), p_ld,
                _stream_ptr(self.device)), "iic_op_layernorm")
        return (out, p) if lora_a_scaled is not None else out

    def op_lora_down(self, x: torch.Tensor, lora_a_scaled: torch.Tensor, p_ld: int = 16) -> torch.Tensor:
        rows, K = x.shape
        p = torch.zeros(rows, p_ld, dtype=self.op_dtype, device=self.device)
        with self._lock, torch.cuda.device(self.device):
            L.check(self.h, self.lib.iic_op_lora_down(self.h, x.data_ptr(), K, rows, lora_a_scaled.data_ptr(),
                                                      lora_a_scaled.shape[1], p.data_ptr(), p_ld,
                                                      _stream_ptr(self.device)), "iic_op_lora_down")
        return p

    def op_attention(self, qkv: torch.Tensor, B: int, T: int, heads: int, impl: int = 0) -> torch.Tensor:
        """impl: 0 = encoder default, 1 = mma.sync kernel, 2 = block-wise tcgen05 kernel (T <= ~760), 3 = whole-row tcgen05 kernel
        (unmasked, T <= 208)"""
        out = torch.empty(B * T, heads * 64, dtype=self.op_dtype, device=self.device)
        with self._lock, torch.cuda.device(self.device):
            L.check(self.h, self.lib.iic_op_attention(self.h, qkv.data_ptr(), out.data_ptr(), B, T, heads, impl,
                                                      _stream_ptr(self.device)), "iic_op_attention")
        return out
