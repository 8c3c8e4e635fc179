"""The knowledge-distillation step with the reference's call surface.

`DistillationStep(student, teacher, kl_weight, share_hidden_states).train_step(batch, temperature)` returns
`(loss, {"loss", "ce_loss", "kl_loss"})` exactly like the closure at ref:training/run_distillation.py:1465-1495, and
`eval_step(batch)` mirrors :1498-1522 (temperature 1).  `loss.backward()` (ref :1609) runs the CUDA backward.

Fused relative to the reference: the two [B, T, V] logits tensors are reduced by ONE kernel (dwb_kd_loss) that
produces CE, KL and d loss / d student_logits (bf16) together -- softmax(t/T), log_softmax(s/T), KLDivLoss, the mask
and the mean (ref :1453-1462, :1484-1493) never exist as tensors.
"""
from __future__ import annotations

import torch

from . import engine, ops


class _KDStepFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, step, batch, temperature, loss_scale, anchor, pre=None):
        student, teacher = step.student, step.teacher
        cfg = student.config
        feats, dec_in, labels = batch["input_features"], batch["decoder_input_ids"], batch["labels"]
        B, T = dec_in.shape
        V = cfg.vocab_size
        if teacher.config.vocab_size != V:
            raise ValueError("student and teacher vocabularies differ")
        train = ctx is not None
        if pre is None:
            pre = step.encode_frozen(feats)
        # ---- student (ref :1472)
        if "student" in pre:
            (enc_s, S), ectx = pre["student"], None
        else:
            enc_s, S, ectx = engine.run_encoder(student, feats, None, save=train)
        sst = engine.state_of(student.model.decoder)
        with engine.nvtx_range("dwb.student_decoder_forward"):
            hf_s, dctx = engine.decoder_forward(sst, dec_in, enc_s, B, S, save=train)
            logits_s = engine.lm_head(sst, hf_s)
        # ---- teacher, no grad (ref :1473-1481)
        tst = engine.state_of(teacher.model.decoder)
        if step.share_hidden_states:
            # teacher(encoder_outputs=student states, labels=labels): its decoder inputs are rebuilt from the labels
            # (HF:models/whisper/modeling_whisper.py:1064-1067), which differs from the student's on prompt-masked rows
            t_in = engine.shift_tokens_right(labels, teacher.config.pad_token_id, teacher.config.decoder_start_token_id)
            enc_t = enc_s
        else:
            t_in = dec_in
            enc_t = pre["teacher"][0]
        with engine.nvtx_range("dwb.teacher_decoder_forward"):
            hf_t, _ = engine.decoder_forward(tst, t_in, enc_t, B, S, save=False)
            logits_t = engine.lm_head(tst, hf_t)
        # ---- fused loss head (ref :1484-1493)
        with engine.nvtx_range("dwb.kd_loss"):
            metrics, dl = ops.kd_loss(logits_s, logits_t, labels, V, temperature, 0.8 * loss_scale, step.kl_weight * loss_scale,
                                      want_grad=train)
        if train:
            ctx.step, ctx.dctx, ctx.ectx, ctx.dl = step, dctx, ectx, dl
        step.last_student_logits = logits_s.view(B, T, -1)[:, :, :V] if step.keep_logits else None
        step.last_teacher_logits = logits_t.view(B, T, -1)[:, :, :V] if step.keep_logits else None
        step.last_encoder_states = enc_s
        return metrics[0].clone(), metrics

    @staticmethod
    def backward(ctx, g_loss, _g_metrics):
        # d loss / d logits was produced by the loss kernel with loss_scale folded in.  The upstream gradient is 1 for
        # `loss.backward()`, 1/gradient_accumulation_steps for accelerator.backward(loss) (ref :1607-1609) or a GradScaler's
        # factor: it is applied on the device (dwb_scale_bf16_dev returns at once when the scalar is 1), never dropped.
        step = ctx.step
        if g_loss is not None:
            ops.scale_bf16_dev(ctx.dl, g_loss.detach().to(torch.float32).reshape(1).contiguous())
        engine.backward_through_model(step.student, ctx.dctx, ctx.ectx, ctx.dl)
        ctx.dctx = ctx.ectx = ctx.dl = None
        return None, None, None, None, None, None


class _PlainCtx:
    """Stand-in for the autograd ctx when forward and backward are issued back to back without autograd (CUDA graph)."""


class DistillationStep:
    """Holds the student / teacher pair the way the reference's closures capture them (ref :1449, :1046-1049)."""

    def __init__(self, student, teacher, kl_weight: float = 1.0, share_hidden_states: bool | None = None, keep_logits=False):
        self.student, self.teacher = student, teacher
        self.kl_weight = float(kl_weight)
        if share_hidden_states is None:      # ref :1046: training_args.freeze_encoder and same d_model
            enc_frozen = not any(p.requires_grad for n, p in student.model.encoder.named_parameters())
            share_hidden_states = enc_frozen and student.config.d_model == teacher.config.d_model
        self.share_hidden_states = bool(share_hidden_states)
        if self.share_hidden_states:         # ref :1047-1049
            teacher.model.encoder = student.model.encoder
        self.keep_logits = keep_logits
        self.last_student_logits = self.last_teacher_logits = self.last_encoder_states = None

    def _anchor(self):
        return next((p for p in self.student.parameters() if p.requires_grad), None)

    @torch.no_grad()
    def encode_frozen(self, input_features):
        """Every encoder forward of the step that takes no gradient: the student's when its encoder is frozen (ref :1023-1026),
        the teacher's own when hidden states are not shared (ref :1481).  None of them reads a trainable weight, so this part
        of step i+1 may run while the optimiser of step i is still updating the student (PipelinedTrainer)."""
        pre = {}
        if not engine.encoder_is_trainable(self.student):
            enc_s, S, _ = engine.run_encoder(self.student, input_features, None, save=False)
            pre["student"] = (enc_s, S)
        if not self.share_hidden_states:
            enc_t, S, _ = engine.run_encoder(self.teacher, input_features, None)
            pre["teacher"] = (enc_t, S)
        return pre

    def train_step(self, batch, temperature: float = 2.0, loss_scale: float = 1.0):
        """loss_scale folds a constant such as 1/gradient_accumulation_steps into the loss kernel for free; a gradient
        supplied to `loss.backward(g)` / accelerator.backward (ref :1607-1609) is honoured as well (applied on the device)."""
        self.student.train()
        self.teacher.eval()
        anchor = self._anchor()
        if anchor is None:
            raise RuntimeError("student has no trainable parameters")
        engine._check_trainable_dtypes(self.student)
        loss, m = _KDStepFn.apply(self, batch, float(temperature), float(loss_scale), anchor, None)
        return loss, {"loss": loss, "ce_loss": m[1], "kl_loss": m[2]}

    @torch.no_grad()
    def forward_backward(self, batch, temperature: float = 2.0, loss_scale: float = 1.0, pre=None):
        """train_step + loss.backward() as one straight-line kernel sequence (no autograd graph): gradients are accumulated
        into .grad exactly as loss.backward() would.  This is what the CUDA-graph wrappers capture.  `pre`: the result of
        encode_frozen() on the same features when the caller issued it separately."""
        self.student.train()
        self.teacher.eval()
        engine._check_trainable_dtypes(self.student)
        ctx = _PlainCtx()
        loss, m = _KDStepFn.forward(ctx, self, batch, float(temperature), float(loss_scale), None, pre)
        engine.backward_through_model(self.student, ctx.dctx, ctx.ectx, ctx.dl)
        return loss, {"loss": loss, "ce_loss": m[1], "kl_loss": m[2]}

    @torch.no_grad()
    def eval_step(self, batch):
        self.student.eval()
        self.teacher.eval()
        _, m = _KDStepFn.forward(None, self, batch, 1.0, 1.0, None, None)
        return {"loss": m[0], "ce_loss": m[1], "kl_loss": m[2]}


def _warm_up(step, static_batch, temperature, loss_scale, warmup):
    """Run the step a few times on a side stream without leaving a trace in the gradients (builds shadows, sets kernel
    attributes and sizes the allocator outside the capture)."""
    params = [p for p in step.student.parameters() if p.requires_grad]
    if any(p.grad is None for p in params):
        raise RuntimeError("bind the gradients to static buffers first (construct optim.FusedAdamW before graph capture)")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        saved = [p.grad.clone() for p in params]
        for _ in range(warmup):
            step.forward_backward(static_batch, temperature, loss_scale)
        for p, g in zip(params, saved):
            p.grad.copy_(g)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    # the bf16 shadows of trainable weights must be re-cast on every replay: invalidate them so that the cast kernels
    # are part of the captured work (frozen weights keep their cached shadows and are not re-cast)
    engine.bump_param_epoch()


class GraphedDistillationStep:
    """`train_step` + `loss.backward()` captured once into a CUDA graph and replayed per batch.

    The KD step issues ~6000 kernel launches with fixed shapes; replaying them as one graph removes the per-launch host
    cost (ctypes call, tensor-map encode, allocator) from the critical path.  Batches are copied into static device
    buffers; parameter gradients land in the same (flat) buffers as in the eager path, so the optimiser and the gradient
    all-reduce are unchanged and stay outside the graph (their scalars change every step)."""

    def __init__(self, step: DistillationStep, example_batch: dict, temperature: float = 2.0, loss_scale: float = 1.0, warmup: int = 2):
        self.step = step
        self.static_batch = {k: v.clone() for k, v in example_batch.items()}
        _warm_up(step, self.static_batch, temperature, loss_scale, warmup)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            loss, metrics = step.forward_backward(self.static_batch, temperature, loss_scale)
            self.loss = loss
            self.metrics = metrics

    def __call__(self, batch: dict | None = None):
        """Copy `batch` (device or pinned host tensors) into the static buffers, replay, return (loss, metrics) tensors that
        are overwritten by the next call.  Gradients are ACCUMULATED into .grad exactly like loss.backward()."""
        if batch is not None:
            for k, dst in self.static_batch.items():
                dst.copy_(batch[k], non_blocking=True)
        self.graph.replay()
        return self.loss, self.metrics


class PipelinedTrainer:
    """The loop body of ref:training/run_distillation.py:1606-1614 -- train_step, backward, (DDP all-reduce), clip_grad_norm_,
    optimizer.step, zero_grad -- as two CUDA graphs and a side stream:

        main stream : [H2D batch i] [graph E_i: gradient-free encoder forwards] ..wait tail i-1.. [graph D_i: decoders, loss, backward]
        side stream :                [all-reduce i-1][clip + AdamW i-1]                            ..wait D_i.. [all-reduce i][AdamW i]

    With the reference recipe (frozen, shared encoder) E is 76 % of the step's FLOPs and reads no trainable weight, so the
    gradient all-reduce and the optimiser of step i-1 run entirely underneath it; with a trainable student encoder E is the
    teacher's encoder forward only.  `gradient_accumulation_steps` = k: the tail runs on every k-th call only (accelerate's
    `accumulate` / no_sync: ONE all-reduce per optimiser step, ref :1607) and 1/k is folded into the loss kernel.
    """

    def __init__(self, step: DistillationStep, optimizer, example_batch: dict, temperature: float = 2.0,
                 gradient_accumulation_steps: int = 1, group=None, warmup: int = 2, tail_grid: int | None = None):
        self.kd, self.opt, self.group = step, optimizer, group
        # optional grid cap of the overlapped clip + AdamW kernels (dwb_set_tail_grid; 0 = none).  Measured at N = 1
        # (profiles/r02_tail_overlap.md): serial tail, overlapped tail and capped grids are within noise of each other (the
        # kernels themselves take 0.85 ms); the side stream pays off through the NCCL all-reduce at N > 1.
        import os
        from . import _abi
        self.tail_grid = int(os.environ.get("DWB_TAIL_GRID", 0 if tail_grid is None else tail_grid))
        self.overlap = os.environ.get("DWB_TAIL_OVERLAP", "1") != "0"      # 0: tail on the main stream (A/B)
        self._abi = _abi
        self.accum = int(gradient_accumulation_steps)
        if self.accum < 1:
            raise ValueError("gradient_accumulation_steps must be >= 1")
        loss_scale = 1.0 / self.accum
        self.static_batch = {k: v.clone() for k, v in example_batch.items()}
        _warm_up(step, self.static_batch, temperature, loss_scale, warmup)
        self.side = torch.cuda.Stream()
        self.bwd_done, self.tail_done = torch.cuda.Event(), torch.cuda.Event()
        self._tail_pending = False
        self._micro = 0
        self.profile = None                   # set to [] to collect (E start, E end, D end, tail start, tail end, D start) CUDA events per step
        self.g_enc = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_enc):
            self.pre = step.encode_frozen(self.static_batch["input_features"])
        if not self.pre:                      # trainable encoder shared with the teacher: nothing is gradient-free
            self.g_enc = None
        self.g_dec = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_dec):
            self.loss, self.metrics = step.forward_backward(self.static_batch, temperature, loss_scale, pre=self.pre)

    def describe(self):
        e = "+".join(sorted(self.pre)) if self.pre else "none"
        return (f"cuda_graph E(encoder fwd: {e}) | cuda_graph D(decoders+loss+bwd); all-reduce+clip+AdamW on a side stream under the next "
                f"step's E; grad_accum {self.accum}")

    def step(self, batch: dict | None = None, tail: bool | None = None):
        """One micro-step.  Returns the loss tensor (overwritten by the next call).  `tail` overrides the accumulation
        schedule (False: never reduce / update; used by bench.py to measure the exposed part of the tail)."""
        main = torch.cuda.current_stream()
        if batch is not None:
            for k, dst in self.static_batch.items():
                dst.copy_(batch[k], non_blocking=True)
        ev = None
        if self.profile is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
            ev[0].record(main)
        if self.g_enc is not None:
            self.g_enc.replay()
        if ev:
            ev[1].record(main)
        if self._tail_pending:                # D reads the weights the previous optimiser step writes and adds into the gradients it zeroes
            main.wait_event(self.tail_done)
            self._tail_pending = False
        if ev:
            ev[5].record(main)                # E end (ev[1]) -> here = time the main stream waited for the previous tail
        self.g_dec.replay()
        if ev:
            ev[2].record(main)
        self._micro += 1
        run_tail = (self._micro % self.accum == 0) if tail is None else tail
        if run_tail and not self.overlap:
            if ev:
                ev[3].record(main)
            self.opt.all_reduce_gradients(self.group)
            self.opt.step()
            if ev:
                ev[4].record(main)
        elif run_tail:
            self.bwd_done.record(main)
            with torch.cuda.stream(self.side):
                self.side.wait_event(self.bwd_done)
                if ev:
                    ev[3].record(self.side)
                with engine.nvtx_range("dwb.grad_all_reduce"):
                    self.opt.all_reduce_gradients(self.group)
                with engine.nvtx_range("dwb.clip_adamw"):
                    self._abi.call("dwb_set_tail_grid", self.tail_grid)      # host-side launch parameter of the next two kernels only
                    try:
                        self.opt.step()
                    finally:
                        self._abi.call("dwb_set_tail_grid", 0)
                if ev:
                    ev[4].record(self.side)
                self.tail_done.record(self.side)
            self._tail_pending = True
        if ev and run_tail:
            self.profile.append(ev)
        return self.loss

    def flush(self):
        """Make the current stream wait for the outstanding all-reduce + optimiser (device-side wait, no host sync)."""
        if self._tail_pending:
            torch.cuda.current_stream().wait_event(self.tail_done)
            self._tail_pending = False
