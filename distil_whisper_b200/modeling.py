"""Whisper model objects with the Hugging Face surface the reference script touches, computed by libdwb.so.

Boundary: ref:training/run_distillation.py builds its student/teacher with
`WhisperForConditionalGeneration.from_pretrained` (:986-1004) and then only uses
  * forward(input_features, decoder_input_ids, labels) / forward(encoder_outputs=..., labels=...) -> .loss .logits
    .encoder_last_hidden_state                                             (:1472-1488, HF:modeling_whisper.py:995-1100)
  * .config.{decoder_start_token_id, d_model, ...}, .model.encoder / .model.decoder / .proj_out as nn.Modules whose
    parameters can be frozen, an encoder that can be re-assigned (:1024-1049), named_parameters()/named_children()
    with nn.LayerNorm children for the weight-decay split (:760-778, :1392-1400), state_dict() names == HF's.
The nn.Linear / nn.LayerNorm / nn.Conv1d / nn.Embedding members below are parameter containers with exactly those
names; their own forward() is never used -- engine.py drives the CUDA kernels on their tensors.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import torch
from torch import nn

from . import engine


@dataclass
class WhisperB200Config:
    """Field names follow transformers.WhisperConfig (HF:models/whisper/configuration_whisper.py)."""
    vocab_size: int = 51866
    num_mel_bins: int = 80
    d_model: int = 1280
    encoder_layers: int = 32
    encoder_attention_heads: int = 20
    encoder_ffn_dim: int = 5120
    decoder_layers: int = 2
    decoder_attention_heads: int = 20
    decoder_ffn_dim: int = 5120
    max_source_positions: int = 1500
    max_target_positions: int = 448
    pad_token_id: int = 50256
    bos_token_id: int = 50257
    eos_token_id: int = 50257
    decoder_start_token_id: int = 50258
    init_std: float = 0.02
    dropout: float = 0.0
    attention_dropout: float = 0.0
    activation_dropout: float = 0.0
    activation_function: str = "gelu"
    scale_embedding: bool = False
    max_length: int = 448
    extra: dict = field(default_factory=dict)

    @classmethod
    def from_any(cls, cfg):
        """Accept a transformers.WhisperConfig, a dict, or an object with the same attribute names."""
        if isinstance(cfg, cls):
            return cfg
        get = (lambda k, d: cfg.get(k, d)) if isinstance(cfg, dict) else (lambda k, d: getattr(cfg, k, d))
        base = cls()
        kw = {f: get(f, getattr(base, f)) for f in cls.__dataclass_fields__ if f != "extra"}
        out = cls(**kw)
        out.validate()
        return out

    def validate(self):
        if self.d_model % self.encoder_attention_heads or self.d_model // self.encoder_attention_heads != 64 \
                or self.d_model // self.decoder_attention_heads != 64:
            raise ValueError("the B200 attention kernels are specialised for head_dim 64 (every Whisper size)")
        if self.activation_function != "gelu":
            raise ValueError("only Whisper's exact-erf GELU is implemented")
        if self.dropout or self.attention_dropout or self.activation_dropout:
            raise ValueError("dropout > 0 is not implemented (Whisper configs use 0.0)")
        if self.scale_embedding:
            raise ValueError("scale_embedding=True is not implemented (Whisper configs use False)")

    def to_dict(self):
        return {f: getattr(self, f) for f in self.__dataclass_fields__ if f != "extra"}


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    """Frozen encoder position table (same closed form as HF:models/whisper/modeling_whisper.py:55-64)."""
    inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2))
    t = torch.arange(length).view(-1, 1) * inv.view(1, -1)
    return torch.cat([t.sin(), t.cos()], dim=1)


@dataclass
class Seq2SeqLMOutputB200:
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    encoder_last_hidden_state: Optional[torch.Tensor] = None
    past_key_values: None = None

    def __getitem__(self, k):
        return getattr(self, k) if isinstance(k, str) else (self.loss, self.logits, self.encoder_last_hidden_state)[k]


@dataclass
class BaseModelOutputB200:
    last_hidden_state: torch.Tensor = None

    def __getitem__(self, i):
        return (self.last_hidden_state,)[i]


class _Attention(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.k_proj = nn.Linear(d, d, bias=False)     # HF:modeling_whisper.py:279 (no key bias)
        self.v_proj = nn.Linear(d, d)
        self.q_proj = nn.Linear(d, d)
        self.out_proj = nn.Linear(d, d)


class _EncoderLayer(nn.Module):
    def __init__(self, d, ffn):
        super().__init__()
        self.self_attn = _Attention(d)
        self.self_attn_layer_norm = nn.LayerNorm(d)
        self.fc1 = nn.Linear(d, ffn)
        self.fc2 = nn.Linear(ffn, d)
        self.final_layer_norm = nn.LayerNorm(d)


class _DecoderLayer(nn.Module):
    def __init__(self, d, ffn):
        super().__init__()
        self.self_attn = _Attention(d)
        self.self_attn_layer_norm = nn.LayerNorm(d)
        self.encoder_attn = _Attention(d)
        self.encoder_attn_layer_norm = nn.LayerNorm(d)
        self.fc1 = nn.Linear(d, ffn)
        self.fc2 = nn.Linear(ffn, d)
        self.final_layer_norm = nn.LayerNorm(d)


class WhisperB200Encoder(nn.Module):
    def __init__(self, config: WhisperB200Config):
        super().__init__()
        self.config = config
        d = config.d_model
        self.conv1 = nn.Conv1d(config.num_mel_bins, d, kernel_size=3, padding=1)
        self.conv2 = nn.Conv1d(d, d, kernel_size=3, stride=2, padding=1)
        self.embed_positions = nn.Embedding(config.max_source_positions, d)
        self.embed_positions.requires_grad_(False)                       # HF:modeling_whisper.py:571
        self.layers = nn.ModuleList([_EncoderLayer(d, config.encoder_ffn_dim) for _ in range(config.encoder_layers)])
        self.layer_norm = nn.LayerNorm(d)
        self.gradient_checkpointing = False

    def forward(self, input_features, **kwargs):
        st = engine.state_of(self)
        enc, _ = engine.encoder_forward(st, input_features, save=False)
        B = input_features.shape[0]
        return BaseModelOutputB200(enc.view(B, -1, self.config.d_model))


class WhisperB200Decoder(nn.Module):
    def __init__(self, config: WhisperB200Config):
        super().__init__()
        self.config = config
        d = config.d_model
        self.embed_tokens = nn.Embedding(config.vocab_size, d, padding_idx=config.pad_token_id)
        self.embed_positions = nn.Embedding(config.max_target_positions, d)
        self.layers = nn.ModuleList([_DecoderLayer(d, config.decoder_ffn_dim) for _ in range(config.decoder_layers)])
        self.layer_norm = nn.LayerNorm(d)
        self.gradient_checkpointing = False


class WhisperB200Model(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.encoder = WhisperB200Encoder(config)
        self.decoder = WhisperB200Decoder(config)


class DistilWhisperB200ForConditionalGeneration(nn.Module):
    """Drop-in for transformers.WhisperForConditionalGeneration on the KD training path."""
    base_model_prefix = "model"

    def __init__(self, config):
        super().__init__()
        self.config = WhisperB200Config.from_any(config)
        self.config.validate()
        self.model = WhisperB200Model(self.config)
        self.proj_out = nn.Linear(self.config.d_model, self.config.vocab_size, bias=False)
        self.proj_out.weight = self.model.decoder.embed_tokens.weight            # tied (HF:modeling_whisper.py:966)
        self.generation_config = None
        self.reset_parameters()

    # ---- init / (de)serialisation -------------------------------------------------------------------------------
    @torch.no_grad()
    def reset_parameters(self):
        std = self.config.init_std
        for m in self.modules():
            if m is self.proj_out:           # tied to embed_tokens: initialised (with its zero padding row) there
                continue
            if isinstance(m, (nn.Linear, nn.Conv1d)):
                m.weight.normal_(0.0, std)
                if m.bias is not None:
                    m.bias.zero_()
            elif isinstance(m, nn.Embedding):
                m.weight.normal_(0.0, std)
                if m.padding_idx is not None:
                    m.weight[m.padding_idx].zero_()
            elif isinstance(m, nn.LayerNorm):
                m.weight.fill_(1.0)
                m.bias.zero_()
        enc = self.model.encoder
        enc.embed_positions.weight.copy_(sinusoids(*enc.embed_positions.weight.shape))

    def load_hf_state_dict(self, sd, strict=True):
        sd = {k: v for k, v in sd.items() if k != "proj_out.weight"}
        missing, unexpected = self.load_state_dict(sd, strict=False)
        missing = [k for k in missing if k != "proj_out.weight"]
        if strict and (missing or unexpected):
            raise RuntimeError(f"state dict mismatch: missing={missing} unexpected={unexpected}")
        engine.invalidate(self)
        return missing, unexpected

    @classmethod
    def from_hf(cls, hf_model, dtype=None):
        """Build from an in-memory transformers.WhisperForConditionalGeneration (weights copied, names identical)."""
        m = cls(hf_model.config)
        m.load_hf_state_dict(hf_model.state_dict())
        m.generation_config = getattr(hf_model, "generation_config", None)
        return m.to(dtype) if dtype is not None else m

    def save_pretrained(self, save_directory, safe_serialization=True, **kwargs):
        """config.json + model.safetensors (or pytorch_model.bin) with Hugging Face's parameter names, loadable by
        transformers.WhisperForConditionalGeneration.from_pretrained and by from_pretrained below (ref:training/
        run_distillation.py:1643-1652, :1754-1760 save the student this way through accelerate)."""
        import json
        import os
        os.makedirs(save_directory, exist_ok=True)
        sd = {k: v.detach().to("cpu").contiguous() for k, v in self.state_dict().items() if k != "proj_out.weight"}
        if safe_serialization:
            from safetensors.torch import save_file
            save_file(sd, os.path.join(save_directory, "model.safetensors"), metadata={"format": "pt"})
        else:
            torch.save(sd, os.path.join(save_directory, "pytorch_model.bin"))
        with open(os.path.join(save_directory, "config.json"), "w") as f:
            json.dump({"model_type": "whisper", "architectures": ["WhisperForConditionalGeneration"], **self.config.to_dict()}, f, indent=2)
        gen = self.generation_config
        if gen is not None:
            gd = gen if isinstance(gen, dict) else (gen.to_dict() if hasattr(gen, "to_dict") else dict(vars(gen)))
            with open(os.path.join(save_directory, "generation_config.json"), "w") as f:
                json.dump(gd, f, indent=2, default=str)

    @classmethod
    def from_pretrained(cls, directory, torch_dtype=None, device=None, **kwargs):
        """Load a directory written by save_pretrained above or by transformers (config.json + model.safetensors /
        pytorch_model.bin, single shard).  Mirrors the call at ref:training/run_distillation.py:986-1004 for local paths."""
        import json
        import os
        with open(os.path.join(directory, "config.json")) as f:
            model = cls(json.load(f))
        st, pt = os.path.join(directory, "model.safetensors"), os.path.join(directory, "pytorch_model.bin")
        if os.path.exists(st):
            from safetensors.torch import load_file
            sd = load_file(st)
        elif os.path.exists(pt):
            sd = torch.load(pt, map_location="cpu")
        else:
            raise FileNotFoundError(f"no model.safetensors / pytorch_model.bin in {directory} (sharded checkpoints are not supported)")
        model.load_hf_state_dict(sd)
        gpath = os.path.join(directory, "generation_config.json")
        if os.path.exists(gpath):
            with open(gpath) as f:
                model.generation_config = json.load(f)
        if torch_dtype is not None:
            model = model.to(torch_dtype)
        return model.to(device) if device is not None else model

    def gradient_checkpointing_enable(self, *a, **k):
        """ref:training/run_distillation.py:1012-1013 calls this when --gradient_checkpointing is set.  This implementation
        keeps every activation of the trainable part resident (they fit in 180 GB at the reference batch size)
        and has no recomputation schedule: the request is acknowledged with a warning, never silently."""
        import warnings
        warnings.warn("distil_whisper_b200: gradient checkpointing is not implemented (activations stay resident in HBM); "
                      "continuing without recomputation", stacklevel=2)
        self.model.encoder.gradient_checkpointing = False
        self.model.decoder.gradient_checkpointing = False

    def get_encoder(self):
        return self.model.encoder

    def get_decoder(self):
        return self.model.decoder

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        engine.invalidate(self)
        return out

    # ---- forward ------------------------------------------------------------------------------------------------
    def forward(self, input_features=None, attention_mask=None, decoder_input_ids=None, decoder_attention_mask=None,
                encoder_outputs=None, past_key_values=None, decoder_inputs_embeds=None, decoder_position_ids=None,
                labels=None, use_cache=None, **kwargs):
        """HF:models/whisper/modeling_whisper.py:995-1100.  Returns an object with .loss / .logits /
        .encoder_last_hidden_state; .logits and .loss carry autograd history (loss.backward() runs the CUDA backward)."""
        if decoder_inputs_embeds is not None or past_key_values is not None or decoder_position_ids is not None:
            raise NotImplementedError("decoder_inputs_embeds / past_key_values / decoder_position_ids are not on the KD training path")
        if labels is not None and decoder_input_ids is None:
            decoder_input_ids = engine.shift_tokens_right(labels, self.config.pad_token_id, self.config.decoder_start_token_id)
        if decoder_input_ids is None:
            raise ValueError("decoder_input_ids or labels are required")
        enc_in = None
        if encoder_outputs is not None:
            enc_in = encoder_outputs[0] if not isinstance(encoder_outputs, torch.Tensor) else encoder_outputs
        logits, enc = engine.ModelForwardFn.run(self, input_features, decoder_input_ids, enc_in)
        loss = None
        if labels is not None:
            loss = engine.CrossEntropyFn.apply(logits, labels, self.config.vocab_size)
        B = decoder_input_ids.shape[0]
        return Seq2SeqLMOutputB200(loss=loss, logits=logits,
                                   encoder_last_hidden_state=enc.view(B, -1, self.config.d_model))

    def generate(self, input_features=None, **kwargs):
        """Greedy decoding with a KV cache for the reference's eval loop (ref:training/run_distillation.py:1428-1446, :1526:
        `student_model.generate(batch["input_features"], max_length=..., num_beams=..., return_timestamps=..., language=...,
        task=...)`) and the pseudo-labelling loop (ref:training/run_pseudo_labelling.py:903).  See generation.py: initial
        tokens from language / task / the generation config, suppress_tokens / begin_suppress_tokens, EOS bookkeeping like
        GenerationMixin; beam search, sampling and timestamp rules raise NotImplementedError."""
        from . import generation
        return generation.generate(self, input_features, **kwargs)
