"""LoRA fine-tuning step of the vision tower on the CUDA engine.

Mirrors the step of /root/reference/train_lora.py:227-252 - symmetric InfoNCE between L2-normalised image and text
features scaled by `logit_scale.exp()`, labels = arange(batch), `loss.backward()`, `clip_grad_norm_(max_norm=1.0)`,
AdamW(lr=1e-4, weight_decay=0.01) over the parameters whose name contains 'lora' - with the LoRA sitting on the
vision MLPs (`mlp.c_fc`, `mlp.c_proj`: what main.py's wrap makes effective) and the text side supplying
fixed target embeddings (BASELINE config 4).  The reference itself fine-tunes the text tower and keeps the vision tower
under no_grad; this is the vision-side counterpart the north star asks for.

Split of work: the engine runs the encoder forward (keeping activations) and the whole backward through the frozen
blocks (dX GEMMs on tcgen05, attention / LayerNorm / GELU backward, LoRA dA/dB reductions); the O(B x width) head + loss
(ln_post, proj, L2-norm, cross-entropy) and the optimizer are ordinary PyTorch, exactly the reference's code path.
Data parallel: per-rank local loss (labels = arange(local batch)), then ONE collective per block - an NCCL all-reduce
(average) of that block's four LoRA gradient tensors (1.47 MB in total at rank 4), issued on a side stream as soon as the
block's backward has been enqueued, so it overlaps the backward of the blocks below.  The gradient-norm clip uses the
averaged (global) gradients.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import _lib as L
from .clip_compat import CLIP, VisionTransformer, _is_lora_wrapped


def _broadcast_lora(params: List[torch.nn.Parameter], process_group=None) -> None:
    """Every rank must start from rank 0's adapters: lora_A comes from the unseeded global RNG (main.py:26), so without this
    the replicas would average gradients but apply them to different parameters and drift apart silently.  One flat
    bucket, one broadcast (what torch DDP does at construction)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1):
        return
    with torch.no_grad():
        flat = torch.cat([p.detach().reshape(-1) for p in params])
        dist.broadcast(flat, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0, group=process_group)
        off = 0
        for p in params:
            p.copy_(flat[off:off + p.numel()].view_as(p))
            off += p.numel()


def _guarded_step(trainer, loss: torch.Tensor) -> float:
    """clip_grad_norm_(1.0) + AdamW step (train_lora.py:249-252), skipped when the gradients are not finite (an fp16 overflow
    of an activation gradient): the step is dropped, the engine's loss scale backs off by 2 and recovers by 2 every 200 good
    steps - one overflow must not poison the AdamW moments and the parameters for good.  One host sync per step (the loss
    read the reference makes anyway), taken before the optimizer launches."""
    total = torch.nn.utils.clip_grad_norm_(trainer.params, max_norm=trainer.max_grad_norm)
    loss_v, total_v = torch.stack([loss.detach().float().reshape(()), total.detach().float().reshape(())]).tolist()
    eng = trainer.eng
    if total_v != total_v or total_v in (float("inf"), float("-inf")):
        trainer.skipped_steps = getattr(trainer, "skipped_steps", 0) + 1
        trainer._good_steps = 0
        eng.loss_scale_backoff = max(getattr(eng, "loss_scale_backoff", 1.0) * 0.5, 2.0 ** -16)
        for p in trainer.params:
            p.grad.zero_()
        return loss_v
    trainer._good_steps = getattr(trainer, "_good_steps", 0) + 1
    if trainer._good_steps % 200 == 0 and getattr(eng, "loss_scale_backoff", 1.0) < 1.0:
        eng.loss_scale_backoff = min(1.0, eng.loss_scale_backoff * 2.0)
    trainer.optimizer.step()
    return loss_v


class VisionLoRATrainer:
    def __init__(self, model: "CLIP | VisionTransformer", lr: float = 1e-4, weight_decay: float = 0.01,
                 max_grad_norm: float = 1.0, logit_scale: Optional[float] = None, process_group=None, overlap: bool = True,
                 distributed: Optional[bool] = None, use_graph: bool = False):
        """distributed: None = data parallel whenever torch.distributed is initialised with more than one rank; False = this
        rank alone (no broadcast, no all-reduce: local gradients).
        use_graph: replay forward + loss + backward of a step as ONE CUDA graph (static shapes and pointers; a power-of-two loss
        scale fixed at capture and re-chosen when the gradient magnitude drifts or a step overflows); the all-reduce of the
        LoRA gradients then runs as one collective on the flat gradient buffer after the replay instead of per block."""
        self.distributed = distributed
        self.visual: VisionTransformer = model.visual if hasattr(model, "visual") else model
        self.logit_scale = float(logit_scale) if logit_scale is not None else (
            float(model.logit_scale.detach().exp()) if hasattr(model, "logit_scale") else 100.0)
        self.max_grad_norm = max_grad_norm
        self.pg = process_group
        self.overlap = overlap
        self.eng = self.visual.engine()
        dev = self.eng.device
        # trainable = LoRA pairs the forward actually uses; an attn.out_proj LoRA never receives a gradient (F4)
        self.slots: List[Tuple[int, int, torch.nn.Module]] = []
        for i, blk in enumerate(self.visual.transformer.resblocks):
            for which, mod in ((L.LORA_C_FC, blk.mlp.c_fc), (L.LORA_C_PROJ, blk.mlp.c_proj)):
                if _is_lora_wrapped(mod):
                    self.slots.append((i, which, mod))
        if not self.slots:
            raise RuntimeError("no LoRA-wrapped mlp.c_fc / mlp.c_proj found: call replace_linears_with_lora(model) first")
        for p in self.visual.parameters():
            p.requires_grad_(False)
        # one flat fp32 gradient bucket per block; each parameter's .grad is a view into it
        self.buckets: Dict[int, torch.Tensor] = {}
        self.params: List[torch.nn.Parameter] = []
        per_layer: Dict[int, List[torch.nn.Parameter]] = {}
        for i, which, mod in self.slots:
            for p in (mod.lora.lora_A, mod.lora.lora_B):
                if p.device != dev or p.dtype != torch.float32:
                    p.data = p.data.to(dev, torch.float32)
                p.requires_grad_(True)
                per_layer.setdefault(i, []).append(p)
                self.params.append(p)
        # ONE flat fp32 gradient buffer; block i's bucket (what its all-reduce moves) is a view of it
        self.flat_grads = torch.zeros(sum(p.numel() for ps in per_layer.values() for p in ps), dtype=torch.float32, device=dev)
        off = 0
        for i, ps in per_layer.items():
            start = off
            for p in ps:
                p.grad = self.flat_grads[off:off + p.numel()].view_as(p)
                off += p.numel()
            self.buckets[i] = self.flat_grads[start:off]
        if distributed is not False:
            _broadcast_lora(self.params, process_group)  # identical adapters on every rank before the optimizer state exists
        # train_lora.py:212 AdamW(lr 1e-4, wd 0.01); fused: one kernel for the 48 small tensors instead of ~15 foreach launches
        self.optimizer = torch.optim.AdamW(self.params, lr=lr, weight_decay=weight_decay, fused=True)
        # CUDA graph of the whole forward + loss + backward (use_graph): ~220 dependent launches per step replayed as one
        self.use_graph = use_graph
        self._graph = None            # (graph, static images, static text, static loss, static |dx| max, loss scale, signature)
        self._graph_steps = 0
        self.comm_stream = torch.cuda.Stream(device=dev) if process_group is not None or self._dist_on() else None
        self._training_weights_sig = None

    @staticmethod
    def _dist_on() -> bool:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    # ------------------------------------------------------------------------------------------------------------
    def _sync(self, refresh: bool = True) -> None:
        """Bring the engine up to date with the parameters (refresh=False: the caller's CUDA graph runs the per-step LoRA
        operand refresh itself).  First call (or after anything but the LoRA VALUES changed):
        full upload.  Every later step: the frozen tensors and all pointers are unchanged, only lora_A / lora_B moved under
        the optimizer, so one `iic_refresh_lora` call rebuilds the derived operands on the device - no host round trip."""
        v = self.visual
        # the tensor OBJECTS are collected once (150 nn.Module attribute walks cost ~0.6 ms of host time per step, which the
        # per-step `float(loss)` turns into GPU idle time); later steps only compare data_ptr / _version of the same objects.
        # Replacing a module or Parameter of the tower after the trainer was built needs a new trainer.
        if getattr(self, "_sd_cache", None) is None:
            self._sd_cache = v._tensors(keep_zero_lora=True)[0]
        sd = self._sd_cache
        sig_w = tuple((k, t.data_ptr(), t._version) for k, t in sd.items())
        sig_p = tuple((i, which, mod.lora.lora_A.data_ptr(), mod.lora.lora_B.data_ptr(), float(mod.lora.scaling),
                       mod.lora.lora_A.grad.data_ptr(), mod.lora.lora_B.grad.data_ptr()) for i, which, mod in self.slots)
        fast = (sig_w, sig_p, id(v._engine), None if v._engine is None else v._engine.op_dtype)
        if v._engine is not None and fast == getattr(self, "_fast_sig", None) and v._sig is not None and v._sig[1] is None:
            if refresh:
                self.eng.refresh_lora()
            return
        eng = v.sync_engine(force=getattr(self, "_fast_sig", None) is not None, keep_zero_lora=True)
        self.eng = eng
        sig = tuple((k, t.data_ptr(), t._version) for k, t in sd.items() if k.endswith(".weight") and "ln_" not in k)
        if sig != self._training_weights_sig:
            eng.enable_training(sd)
            self._training_weights_sig = sig
        for i, which, mod in self.slots:
            lo = mod.lora
            eng.set_lora_train(i, which, lo.lora_A, lo.lora_B, float(lo.scaling), lo.lora_A.grad, lo.lora_B.grad)
            eng.set_lora_source(i, which, lo.lora_A, lo.lora_B, float(lo.scaling))
        self._fast_sig = (sig_w, sig_p, id(eng), eng.op_dtype)
        # the engine's LoRA operands are from now on refreshed behind sync_engine's back: forget its LoRA signature so that
        # an inference call on the same model re-uploads them
        v._sig = (v._sig[0], None)

    def head_and_loss(self, x_cls: torch.Tensor, text_features: torch.Tensor) -> torch.Tensor:
        """ln_post -> proj -> L2 -> symmetric InfoNCE (train_lora.py:241-246), fp32."""
        v = self.visual
        f = F.layer_norm(x_cls, (x_cls.shape[-1],), v.ln_post.weight.detach().float(), v.ln_post.bias.detach().float(), 1e-5)
        f = f @ v.proj.detach().float()
        f = f / f.norm(dim=-1, keepdim=True)
        logits_per_image = (f @ text_features.t()) * self.logit_scale
        labels = torch.arange(x_cls.shape[0], device=x_cls.device)
        return (F.cross_entropy(logits_per_image, labels) + F.cross_entropy(logits_per_image.t(), labels)) / 2

    # ------------------------------------------------------------------------------------------------------------
    def _fb_body(self, images: torch.Tensor, text_features: torch.Tensor, loss_scale, layer_done=None):
        """preprocess -> forward (activations kept) -> head + loss (PyTorch, O(B x width)) -> backward; returns (loss, max |dx_cls|)"""
        eng = self.eng
        if images.dtype == torch.uint8:
            patches, B = eng.preprocess_same_size(images), images.shape[0]
        else:
            patches, B = eng.patchify(images), images.shape[0]
        x_cls = eng.train_forward(patches, B).requires_grad_(True)
        loss = self.head_and_loss(x_cls, text_features)
        (dx_cls,) = torch.autograd.grad(loss, x_cls)
        amax = dx_cls.abs().max()
        eng.train_backward(dx_cls, layer_done=layer_done, loss_scale=loss_scale)
        return loss.detach(), amax

    def _forward_backward_graphed(self, images: torch.Tensor, text_features: torch.Tensor) -> torch.Tensor:
        import math
        import torch.distributed as dist
        eng = self.eng
        dev = eng.device
        sig = (tuple(images.shape), images.dtype, tuple(text_features.shape), getattr(self, "_fast_sig", None) is not None,
               id(eng), float(getattr(eng, "loss_scale_backoff", 1.0)))
        g = self._graph
        if g is not None and g["sig"] != sig:
            g = self._graph = None
        if g is not None and self._graph_steps % 64 == 63:
            # the fixed loss scale follows the gradient magnitude: one 4-byte read every 64 steps, re-capture on a drift of 2^3
            amax = float(g["amax"])
            want = 2.0 ** math.floor(math.log2(256.0 / amax)) if amax > 0 and math.isfinite(amax) else g["scale"]
            if not (g["scale"] / 8.0 <= want * sig[-1] <= g["scale"] * 8.0):
                g = self._graph = None
        if g is None:
            with torch.cuda.device(dev):
                s_img = torch.empty_like(images, device=dev)
                s_txt = torch.empty(text_features.shape, dtype=torch.float32, device=dev)
                s_img.copy_(images)
                s_txt.copy_(text_features)
                # eager warm-up step on a side stream: one-off kernel attributes, workspace allocation, autograd buffers, and the
                # gradient magnitude that picks the loss scale
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    eng.refresh_lora()
                    _, amax0 = self._fb_body(s_img, s_txt, 1.0)
                torch.cuda.current_stream(dev).wait_stream(side)
                a0 = float(amax0)
                scale = (2.0 ** math.floor(math.log2(256.0 / a0)) if a0 > 0 and math.isfinite(a0) else 1.0) * sig[-1]
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    eng.refresh_lora()          # derived LoRA operands from the parameters the optimizer just moved
                    loss, amax = self._fb_body(s_img, s_txt, scale)
            g = self._graph = {"graph": graph, "img": s_img, "txt": s_txt, "loss": loss, "amax": amax, "scale": scale, "sig": sig}
            self._graph_steps = 0
        g["img"].copy_(images, non_blocking=True)
        g["txt"].copy_(text_features, non_blocking=True)
        g["graph"].replay()
        self._graph_steps += 1
        if self._dist_on() and self.distributed is not False:
            dist.all_reduce(self.flat_grads, op=dist.ReduceOp.AVG, group=self.pg)      # 1.5 MB: one collective after the replay
        return g["loss"]

    def forward_backward(self, images: torch.Tensor, text_features: torch.Tensor) -> torch.Tensor:
        """One forward + backward; LoRA gradients land in the parameters' .grad (averaged over ranks when distributed)."""
        import torch.distributed as dist
        self._sync(refresh=not (self.use_graph and self._graph is not None))
        eng = self.eng
        if self.use_graph and getattr(self, "_fast_sig", None) is not None:
            return self._forward_backward_graphed(images.to(eng.device), text_features.detach().to(eng.device, torch.float32))
        if images.dtype == torch.uint8:
            patches, B = eng.preprocess_same_size(images.to(eng.device)), images.shape[0]
        else:
            patches, B = eng.patchify(images.to(eng.device)), images.shape[0]
        x_cls = eng.train_forward(patches, B).requires_grad_(True)
        loss = self.head_and_loss(x_cls, text_features.detach().to(eng.device, torch.float32))
        (dx_cls,) = torch.autograd.grad(loss, x_cls)
        works = []
        distributed = self._dist_on() and self.distributed is not False

        def layer_done(layer: int) -> None:
            if not distributed or layer not in self.buckets:
                return
            if self.overlap:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(eng.device))
                with torch.cuda.stream(self.comm_stream):
                    self.comm_stream.wait_event(ev)
                    works.append(dist.all_reduce(self.buckets[layer], op=dist.ReduceOp.AVG, group=self.pg, async_op=True))
            else:
                works.append(dist.all_reduce(self.buckets[layer], op=dist.ReduceOp.AVG, group=self.pg, async_op=True))

        eng.train_backward(dx_cls, layer_done=layer_done)
        for w in works:
            w.wait()
        if distributed and self.overlap:
            torch.cuda.current_stream(eng.device).wait_stream(self.comm_stream)
        return loss.detach()

    def step(self, images: torch.Tensor, text_features: torch.Tensor) -> float:
        """train_lora.py:249-252: backward, clip_grad_norm_(1.0), optimizer.step()."""
        loss = self.forward_backward(images, text_features)
        return _guarded_step(self, loss)


class TextLoRATrainer:
    """The step /root/reference/train_lora.py:231-252 actually runs: image features from the frozen vision tower under
    no_grad, text features through the LoRA on the TEXT tower's MLPs (`_replace_text_linears_with_lora`, train_lora.py:62-100;
    its attn.out_proj pairs are dead in nn.MultiheadAttention's forward and never receive a gradient), symmetric InfoNCE with
    `logit_scale.exp()`, backward, clip_grad_norm_(1.0), AdamW(lr 1e-4, wd 0.01) over the parameters named '*lora*'.

    The text tower runs on a sequence engine (77 tokens, causal attention): forward keeping activations, backward through
    the frozen blocks with the causal tcgen05 attention backward, LoRA gradients by the same kernels as the vision trainer.
    token_embedding + positional_embedding in front and ln_final / text_projection / loss behind stay in PyTorch (O(B x width)).
    Data parallel exactly like VisionLoRATrainer: one all-reduce (average) per block, overlapped with the blocks below."""

    def __init__(self, model: CLIP, lr: float = 1e-4, weight_decay: float = 0.01, max_grad_norm: float = 1.0,
                 logit_scale: Optional[float] = None, process_group=None, overlap: bool = True):
        self.model = model
        self.logit_scale = float(logit_scale) if logit_scale is not None else float(model.logit_scale.detach().exp())
        self.max_grad_norm = max_grad_norm
        self.pg = process_group
        self.overlap = overlap
        dev = model.token_embedding.weight.device
        if dev.type != "cuda":
            raise RuntimeError("TextLoRATrainer needs the model on a CUDA device (B200); there is no CPU fallback")
        self.slots: List[Tuple[int, int, torch.nn.Module]] = []
        for i, blk in enumerate(model.transformer.resblocks):
            for which, mod in ((L.LORA_C_FC, blk.mlp.c_fc), (L.LORA_C_PROJ, blk.mlp.c_proj)):
                if _is_lora_wrapped(mod):
                    self.slots.append((i, which, mod))
        if not self.slots:
            raise RuntimeError("no LoRA-wrapped text mlp.c_fc / mlp.c_proj found: wrap the text tower first")
        for p in model.parameters():
            p.requires_grad_(False)
        self.buckets: Dict[int, torch.Tensor] = {}
        self.params: List[torch.nn.Parameter] = []
        per_layer: Dict[int, List[torch.nn.Parameter]] = {}
        for i, which, mod in self.slots:
            for p in (mod.lora.lora_A, mod.lora.lora_B):
                if p.device != dev or p.dtype != torch.float32:
                    p.data = p.data.to(dev, torch.float32)
                p.requires_grad_(True)
                per_layer.setdefault(i, []).append(p)
                self.params.append(p)
        for i, ps in per_layer.items():
            flat = torch.zeros(sum(p.numel() for p in ps), dtype=torch.float32, device=dev)
            off = 0
            for p in ps:
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
            self.buckets[i] = flat
        _broadcast_lora(self.params, process_group)
        self.optimizer = torch.optim.AdamW(self.params, lr=lr, weight_decay=weight_decay)   # train_lora.py:212
        self.comm_stream = torch.cuda.Stream(device=dev) if process_group is not None or VisionLoRATrainer._dist_on() else None
        self.eng = None
        self._sig = None

    def _sync(self) -> None:
        m = self.model
        if self.eng is None:
            self.eng = m.sync_text_engine(force=True)
            m._text_sig = None   # inference through model.encode_text re-uploads what the optimizer moved
        eng = self.eng
        if getattr(self, "_sd_cache", None) is None:
            sd: Dict[str, torch.Tensor] = {}
            for i, blk in enumerate(m.transformer.resblocks):
                p = f"transformer.resblocks.{i}."
                sd[p + "attn.in_proj_weight"] = blk.attn.in_proj_weight
                for name, mod in (("attn.out_proj", blk.attn.out_proj), ("mlp.c_fc", blk.mlp.c_fc), ("mlp.c_proj", blk.mlp.c_proj)):
                    sd[p + name + ".weight"] = mod.weight
            self._sd_cache = sd
        sd = self._sd_cache
        sig = (tuple((k, t.data_ptr(), t._version) for k, t in sd.items()),
               tuple((i, which, mod.lora.lora_A.data_ptr(), mod.lora.lora_B.data_ptr(), float(mod.lora.scaling),
                      mod.lora.lora_A.grad.data_ptr(), mod.lora.lora_B.grad.data_ptr()) for i, which, mod in self.slots))
        if sig == self._sig:
            eng.refresh_lora()
            return
        eng.enable_training(sd)
        for i, which, mod in self.slots:
            lo = mod.lora
            eng.set_lora(i, which, lo.lora_A, lo.lora_B, float(lo.scaling))      # also pairs whose B is still zero
            eng.set_lora_train(i, which, lo.lora_A, lo.lora_B, float(lo.scaling), lo.lora_A.grad, lo.lora_B.grad)
            eng.set_lora_source(i, which, lo.lora_A, lo.lora_B, float(lo.scaling))
        self._sig = sig
        m._text_sig = None

    def head_and_loss(self, x_eot: torch.Tensor, image_features: torch.Tensor) -> torch.Tensor:
        """ln_final -> @ text_projection -> L2 -> symmetric InfoNCE (train_lora.py:236-246), fp32."""
        m = self.model
        f = F.layer_norm(x_eot, (x_eot.shape[-1],), m.ln_final.weight.detach().float(), m.ln_final.bias.detach().float(), 1e-5)
        f = f @ m.text_projection.detach().float()
        f = f / f.norm(dim=-1, keepdim=True)
        logits_per_image = (image_features @ f.t()) * self.logit_scale
        labels = torch.arange(x_eot.shape[0], device=x_eot.device)
        return (F.cross_entropy(logits_per_image, labels) + F.cross_entropy(logits_per_image.t(), labels)) / 2

    def forward_backward(self, image_features: torch.Tensor, tokens: torch.Tensor) -> torch.Tensor:
        """image_features [B, E] (L2-normalised, no gradient: train_lora.py:232-234), tokens [B, 77]."""
        import torch.distributed as dist
        self._sync()
        eng, m = self.eng, self.model
        tokens = tokens.to(eng.device)
        with torch.no_grad():
            x = m.token_embedding(tokens).float() + m.positional_embedding.float()
        eot = tokens.argmax(dim=-1)
        x_eot = eng.train_forward_sequence(x, eot).requires_grad_(True)
        loss = self.head_and_loss(x_eot, image_features.detach().to(eng.device, torch.float32))
        (dx,) = torch.autograd.grad(loss, x_eot)
        works = []
        distributed = VisionLoRATrainer._dist_on()

        def layer_done(layer: int) -> None:
            if not distributed or layer not in self.buckets:
                return
            if self.overlap:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(eng.device))
                with torch.cuda.stream(self.comm_stream):
                    self.comm_stream.wait_event(ev)
                    works.append(dist.all_reduce(self.buckets[layer], op=dist.ReduceOp.AVG, group=self.pg, async_op=True))
            else:
                works.append(dist.all_reduce(self.buckets[layer], op=dist.ReduceOp.AVG, group=self.pg, async_op=True))

        eng.train_backward(dx, layer_done=layer_done, row_index=eot)
        for w in works:
            w.wait()
        if distributed and self.overlap:
            torch.cuda.current_stream(eng.device).wait_stream(self.comm_stream)
        return loss.detach()

    def step(self, images_or_features: torch.Tensor, tokens: torch.Tensor) -> float:
        """One optimisation step.  `images_or_features`: a preprocessed image batch [B,3,R,R] (encoded by the frozen vision
        tower on its engine under no_grad, as train_lora.py:232-234 does) or precomputed L2-normalised image features [B, E]."""
        t = images_or_features
        if t.dim() == 4:
            with torch.no_grad():
                f = self.model.encode_image(t).float()
                t = f / f.norm(dim=-1, keepdim=True)
        loss = self.forward_backward(t, tokens)
        return _guarded_step(self, loss)
