"""The reference's own CPU implementation of the KD step, for bench.py's `cpu_baseline` / `--impl reference` legs.
TEST / BENCH INFRASTRUCTURE ONLY -- never imported by distil_whisper_b200.

kind == "reference": Hugging Face `WhisperForConditionalGeneration` modules (the code the reference script delegates to,
ref:training/run_distillation.py:986-1004) driven by a literal restatement of train_step (ref :1453-1495) and of the
loop body (ref :1606-1614: backward, clip_grad_norm_, AdamW.step, zero_grad), on the host cores in fp32.
kind == "port": the same step on oracle/whisper_oracle.py when `transformers` is not importable on the box.
(The script itself cannot run here or on the GPU box: `accelerate` and `evaluate` are not installed)
"""
from __future__ import annotations

import time

import torch
from torch import nn

from . import whisper_oracle as wo


def _hf_available():
    try:
        import transformers  # noqa: F401
        from transformers import WhisperConfig, WhisperForConditionalGeneration  # noqa: F401
        return True
    except Exception:  # noqa: BLE001
        return False


def _hf_model(dims: wo.WhisperDims, dtype=torch.float32):
    from transformers import WhisperConfig, WhisperForConditionalGeneration
    cfg = WhisperConfig(
        vocab_size=dims.vocab_size, num_mel_bins=dims.num_mel_bins, d_model=dims.d_model,
        encoder_layers=dims.encoder_layers, encoder_attention_heads=dims.encoder_attention_heads,
        encoder_ffn_dim=dims.encoder_ffn_dim, decoder_layers=dims.decoder_layers,
        decoder_attention_heads=dims.decoder_attention_heads, decoder_ffn_dim=dims.decoder_ffn_dim,
        max_source_positions=dims.max_source_positions, max_target_positions=dims.max_target_positions,
        pad_token_id=dims.pad_token_id, bos_token_id=dims.pad_token_id, eos_token_id=dims.pad_token_id,
        decoder_start_token_id=dims.decoder_start_token_id, suppress_tokens=None, begin_suppress_tokens=None)
    return WhisperForConditionalGeneration(cfg).to(dtype)


class ReferenceKDStep:
    """Frozen + shared encoder recipe (ref README `--freeze_encoder`) or full (variant A) on the CPU."""

    def __init__(self, student_dims, teacher_dims, freeze_encoder=True, threads=None, lr=1e-4, max_grad_norm=1.0, kl_weight=1.0,
                 device="cpu"):
        """device="cuda": the reference's GPU configuration for context (fp32 student under bf16 autocast, bf16 teacher, sdpa,
        ref:training/run_distillation.py:798-813,985-1004) -- needs transformers."""
        self.device = torch.device(device)
        if threads:
            torch.set_num_threads(threads)
        self.cores = torch.get_num_threads()
        self.kind = "reference" if _hf_available() else "port"
        self.freeze_encoder = freeze_encoder
        self.kl_weight, self.max_grad_norm = kl_weight, max_grad_norm
        self.sc, self.tc = student_dims, teacher_dims
        if self.kind == "reference":
            self.student = _hf_model(student_dims).to(self.device)
            self.teacher = _hf_model(teacher_dims, torch.bfloat16 if self.device.type == "cuda" else torch.float32).to(self.device)
            if freeze_encoder:
                for p in self.student.model.encoder.parameters():
                    p.requires_grad = False
                self.teacher.model.encoder = self.student.model.encoder          # ref :1046-1049
            params = [p for p in self.student.parameters() if p.requires_grad]
        else:
            self.ssd = wo.init_state_dict(student_dims, 1, perturb=False)
            self.tsd = wo.init_state_dict(teacher_dims, 2, perturb=False)
            for k, v in self.ssd.items():
                v.requires_grad_(not (freeze_encoder and k.startswith("model.encoder.")) and k != "model.encoder.embed_positions.weight")
            if freeze_encoder:
                for k in list(self.tsd):
                    if k.startswith("model.encoder."):
                        self.tsd[k] = self.ssd[k]
            params = [v for v in self.ssd.values() if v.requires_grad]
        self.params = params
        self.opt = torch.optim.AdamW(params, lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)

    def step(self, batch, temperature=2.0):
        if self.kind == "reference":
            from transformers.modeling_outputs import BaseModelOutput
            self.student.train()
            self.teacher.eval()
            cuda = self.device.type == "cuda"
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=cuda):
                so = self.student(**batch)
            with torch.no_grad():
                if self.freeze_encoder:
                    enc = so.encoder_last_hidden_state.to(torch.bfloat16) if cuda else so.encoder_last_hidden_state
                    to = self.teacher(encoder_outputs=BaseModelOutput(enc), labels=batch["labels"])
                else:
                    tb = {k: (v.to(torch.bfloat16) if (cuda and v.is_floating_point()) else v) for k, v in batch.items()}
                    to = self.teacher(**tb)
            so.logits = so.logits.float()
            to.logits = to.logits.float()
            ce = so.loss
            td = nn.functional.softmax(to.logits / temperature, dim=-1)
            sd = nn.functional.log_softmax(so.logits / temperature, dim=-1)
            kl = wo.kl_divergence(td, sd, batch["labels"]) * temperature ** 2
            loss = 0.8 * ce + self.kl_weight * kl
        else:
            loss, _, _, _ = wo.kd_train_step(self.ssd, self.sc, self.tsd, self.tc, batch, temperature, self.kl_weight,
                                             share_hidden_states=self.freeze_encoder)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.params, self.max_grad_norm)
        self.opt.step()
        self.opt.zero_grad()
        return float(loss.detach())

    def time_steps(self, batch, steps, warmup):
        cuda = self.device.type == "cuda"
        for _ in range(warmup):
            self.step(batch)
        if cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        last = None
        for _ in range(steps):
            last = self.step(batch)
        if cuda:
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return dt / steps, last
