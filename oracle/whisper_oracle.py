"""CPU oracle for the Distil-Whisper KD training step.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this file.  The product path (distil_whisper_b200/) never does; it fails loudly when the CUDA
extension is missing.

This is a plain PyTorch fp32 restatement of the arithmetic the reference reaches on its hot path:

  * ref:training/run_distillation.py:1453-1462   kl_divergence
  * ref:training/run_distillation.py:1465-1495   train_step (student fwd, teacher fwd, CE + KL mix)
  * ref:training/run_distillation.py:404-478     collator label/decoder-input construction
  * HF:models/whisper/modeling_whisper.py:55-64  sinusoids
  * HF:models/whisper/modeling_whisper.py:68-81  shift_tokens_right
  * HF:models/whisper/modeling_whisper.py:284-357 WhisperAttention (q pre-scaled, k no bias)
  * HF:models/whisper/modeling_whisper.py:380-414 encoder layer, :449-506 decoder layer
  * HF:models/whisper/modeling_whisper.py:593-647 encoder, :691-796 decoder
  * HF:models/whisper/modeling_whisper.py:995-1100 LM head (tied, no bias) + CrossEntropyLoss(-100)

(HF = transformers 5.5.0 as installed in the build container; the reference only pins
transformers>=4.35.1 and has no tests or golden vectors of its own)

Pinning: tests/golden/kd_tiny.npz was produced by oracle/gen_golden.py from the *installed HF
classes* driven by a literal copy of train_step; tests/test_oracle.py checks this restatement
against it, so the oracle is pinned to HF 5.5.0 outputs, not to itself.

State dicts use the HF parameter names verbatim (model.encoder.layers.N.self_attn.q_proj.weight ...).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class WhisperDims:
    """Shape-only config; field names follow HF WhisperConfig."""
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
    decoder_start_token_id: int = 50258

    def to_dict(self):
        return asdict(self)


# Dims are not in the reference; they reproduce README parameter counts (SURVEY.md "verified").
PRESETS = {
    "distil-large-v3": WhisperDims(),
    "large-v3": WhisperDims(decoder_layers=32),
    "distil-small.en": WhisperDims(vocab_size=51864, d_model=768, encoder_layers=12, encoder_attention_heads=12,
                                   encoder_ffn_dim=3072, decoder_layers=4, decoder_attention_heads=12,
                                   decoder_ffn_dim=3072, pad_token_id=50256, decoder_start_token_id=50257),
    "small.en-teacher": WhisperDims(vocab_size=51864, d_model=768, encoder_layers=12, encoder_attention_heads=12,
                                    encoder_ffn_dim=3072, decoder_layers=12, decoder_attention_heads=12,
                                    decoder_ffn_dim=3072, pad_token_id=50256, decoder_start_token_id=50257),
    "distil-medium.en": WhisperDims(vocab_size=51864, d_model=1024, encoder_layers=24, encoder_attention_heads=16,
                                    encoder_ffn_dim=4096, decoder_layers=2, decoder_attention_heads=16,
                                    decoder_ffn_dim=4096, pad_token_id=50256, decoder_start_token_id=50257),
    # tiny shapes used by the golden fixtures / CPU tests (head_dim stays 64 like every Whisper size)
    "tiny-student": WhisperDims(vocab_size=515, d_model=128, encoder_layers=2, encoder_attention_heads=2,
                                encoder_ffn_dim=256, decoder_layers=1, decoder_attention_heads=2,
                                decoder_ffn_dim=256, max_source_positions=50, max_target_positions=32,
                                pad_token_id=500, decoder_start_token_id=501),
    "tiny-teacher": WhisperDims(vocab_size=515, d_model=128, encoder_layers=2, encoder_attention_heads=2,
                                encoder_ffn_dim=256, decoder_layers=3, decoder_attention_heads=2,
                                decoder_ffn_dim=256, max_source_positions=50, max_target_positions=32,
                                pad_token_id=500, decoder_start_token_id=501),
}


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    """HF:models/whisper/modeling_whisper.py:55-64."""
    log_inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-log_inc * torch.arange(channels // 2))
    t = torch.arange(length).view(-1, 1) * inv.view(1, -1)
    return torch.cat([t.sin(), t.cos()], dim=1)


def param_shapes(c: WhisperDims) -> dict:
    """HF state-dict names -> shapes (tied proj_out.weight omitted: it IS decoder.embed_tokens.weight)."""
    d = c.d_model
    s = {}
    s["model.encoder.conv1.weight"] = (d, c.num_mel_bins, 3)
    s["model.encoder.conv1.bias"] = (d,)
    s["model.encoder.conv2.weight"] = (d, d, 3)
    s["model.encoder.conv2.bias"] = (d,)
    s["model.encoder.embed_positions.weight"] = (c.max_source_positions, d)

    def attn(prefix):
        s[prefix + ".k_proj.weight"] = (d, d)
        for n in ("v_proj", "q_proj", "out_proj"):
            s[f"{prefix}.{n}.weight"] = (d, d)
            s[f"{prefix}.{n}.bias"] = (d,)

    def ln(prefix):
        s[prefix + ".weight"] = (d,)
        s[prefix + ".bias"] = (d,)

    for i in range(c.encoder_layers):
        p = f"model.encoder.layers.{i}"
        attn(p + ".self_attn")
        ln(p + ".self_attn_layer_norm")
        s[p + ".fc1.weight"] = (c.encoder_ffn_dim, d)
        s[p + ".fc1.bias"] = (c.encoder_ffn_dim,)
        s[p + ".fc2.weight"] = (d, c.encoder_ffn_dim)
        s[p + ".fc2.bias"] = (d,)
        ln(p + ".final_layer_norm")
    ln("model.encoder.layer_norm")
    s["model.decoder.embed_tokens.weight"] = (c.vocab_size, d)
    s["model.decoder.embed_positions.weight"] = (c.max_target_positions, d)
    for i in range(c.decoder_layers):
        p = f"model.decoder.layers.{i}"
        attn(p + ".self_attn")
        ln(p + ".self_attn_layer_norm")
        attn(p + ".encoder_attn")
        ln(p + ".encoder_attn_layer_norm")
        s[p + ".fc1.weight"] = (c.decoder_ffn_dim, d)
        s[p + ".fc1.bias"] = (c.decoder_ffn_dim,)
        s[p + ".fc2.weight"] = (d, c.decoder_ffn_dim)
        s[p + ".fc2.bias"] = (d,)
        ln(p + ".final_layer_norm")
    ln("model.decoder.layer_norm")
    return s


def init_state_dict(c: WhisperDims, seed: int, std: float = 0.02, perturb: bool = True) -> dict:
    """Deterministic weights from numpy's frozen legacy RandomState (stable across numpy versions),
    so that fixtures need not carry weights.  Distribution follows HF `_init_weights` (normal(0, init_std)
    for Linear/Conv/Embedding weights, sinusoid encoder positions, HF:modeling_whisper.py `_init_weights`)
    except that, with perturb=True, biases and LayerNorm affine parameters are also randomised so that
    every term of the arithmetic is exercised by the parity tests."""
    rs = np.random.RandomState(seed)
    sd = {}
    for name, shape in param_shapes(c).items():
        if name == "model.encoder.embed_positions.weight":
            sd[name] = sinusoids(*shape).float()
            continue
        is_ln = "layer_norm" in name
        if is_ln and name.endswith("weight"):
            w = 1.0 + (0.1 * rs.randn(*shape) if perturb else 0.0)
        elif name.endswith("bias"):
            w = (std * rs.randn(*shape)) if perturb else np.zeros(shape)
        else:
            w = std * rs.randn(*shape)
        sd[name] = torch.from_numpy(np.asarray(w, dtype=np.float32).reshape(shape)).contiguous()
    # nn.Embedding(padding_idx=pad) zeroes that row at init (HF:modeling_whisper.py:676)
    sd["model.decoder.embed_tokens.weight"][c.pad_token_id].zero_()
    return sd


def shift_tokens_right(input_ids: torch.Tensor, pad_token_id: int, decoder_start_token_id: int) -> torch.Tensor:
    """HF:models/whisper/modeling_whisper.py:68-81."""
    out = input_ids.new_zeros(input_ids.shape)
    out[:, 1:] = input_ids[:, :-1].clone()
    out[:, 0] = decoder_start_token_id
    out.masked_fill_(out == -100, pad_token_id)
    return out


def _attention(sd, prefix, n_heads, x, kv_src, causal):
    """HF:models/whisper/modeling_whisper.py:284-357 with the sdpa interface (scaling=1.0, q pre-scaled)."""
    B, Tq, d = x.shape
    hd = d // n_heads
    q = F.linear(x, sd[prefix + ".q_proj.weight"], sd[prefix + ".q_proj.bias"]) * (hd ** -0.5)
    k = F.linear(kv_src, sd[prefix + ".k_proj.weight"])
    v = F.linear(kv_src, sd[prefix + ".v_proj.weight"], sd[prefix + ".v_proj.bias"])
    q = q.view(B, Tq, n_heads, hd).transpose(1, 2)
    k = k.view(B, -1, n_heads, hd).transpose(1, 2)
    v = v.view(B, -1, n_heads, hd).transpose(1, 2)
    s = q @ k.transpose(-1, -2)
    if causal:
        Tk = k.shape[2]
        mask = torch.ones(Tq, Tk, dtype=torch.bool, device=x.device).tril()
        s = s.masked_fill(~mask, float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, Tq, d)
    return F.linear(o, sd[prefix + ".out_proj.weight"], sd[prefix + ".out_proj.bias"])


def _ln(sd, prefix, x):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-5)


def encoder_forward(sd: dict, c: WhisperDims, input_features: torch.Tensor) -> torch.Tensor:
    """HF:models/whisper/modeling_whisper.py:593-647.  input_features [B, n_mels, 2*max_source_positions]."""
    expected = 2 * c.max_source_positions
    if input_features.shape[-1] != expected:
        raise ValueError(f"Whisper expects the mel input features to be of length {expected}, "
                         f"but found {input_features.shape[-1]}.")
    x = F.gelu(F.conv1d(input_features, sd["model.encoder.conv1.weight"], sd["model.encoder.conv1.bias"], padding=1))
    x = F.gelu(F.conv1d(x, sd["model.encoder.conv2.weight"], sd["model.encoder.conv2.bias"], stride=2, padding=1))
    x = x.permute(0, 2, 1)
    x = x + sd["model.encoder.embed_positions.weight"][: x.shape[1]]
    for i in range(c.encoder_layers):
        p = f"model.encoder.layers.{i}"
        h = _ln(sd, p + ".self_attn_layer_norm", x)
        x = x + _attention(sd, p + ".self_attn", c.encoder_attention_heads, h, h, causal=False)
        h = _ln(sd, p + ".final_layer_norm", x)
        h = F.gelu(F.linear(h, sd[p + ".fc1.weight"], sd[p + ".fc1.bias"]))
        x = x + F.linear(h, sd[p + ".fc2.weight"], sd[p + ".fc2.bias"])
    return _ln(sd, "model.encoder.layer_norm", x)


def decoder_forward(sd: dict, c: WhisperDims, decoder_input_ids: torch.Tensor, enc: torch.Tensor) -> torch.Tensor:
    """HF:models/whisper/modeling_whisper.py:691-796 (training call: no cache, positions 0..T-1)."""
    T = decoder_input_ids.shape[1]
    x = F.embedding(decoder_input_ids, sd["model.decoder.embed_tokens.weight"], padding_idx=c.pad_token_id)
    x = x + sd["model.decoder.embed_positions.weight"][:T]
    for i in range(c.decoder_layers):
        p = f"model.decoder.layers.{i}"
        h = _ln(sd, p + ".self_attn_layer_norm", x)
        x = x + _attention(sd, p + ".self_attn", c.decoder_attention_heads, h, h, causal=True)
        h = _ln(sd, p + ".encoder_attn_layer_norm", x)
        x = x + _attention(sd, p + ".encoder_attn", c.decoder_attention_heads, h, enc, causal=False)
        h = _ln(sd, p + ".final_layer_norm", x)
        h = F.gelu(F.linear(h, sd[p + ".fc1.weight"], sd[p + ".fc1.bias"]))
        x = x + F.linear(h, sd[p + ".fc2.weight"], sd[p + ".fc2.bias"])
    return _ln(sd, "model.decoder.layer_norm", x)


def model_forward(sd: dict, c: WhisperDims, input_features=None, decoder_input_ids=None, labels=None,
                  encoder_hidden_states=None):
    """WhisperForConditionalGeneration.forward, HF:models/whisper/modeling_whisper.py:995-1100.
    Returns dict(loss, logits, encoder_last_hidden_state)."""
    if labels is not None and decoder_input_ids is None:
        decoder_input_ids = shift_tokens_right(labels, c.pad_token_id, c.decoder_start_token_id)   # :1064-1067
    enc = encoder_hidden_states if encoder_hidden_states is not None else encoder_forward(sd, c, input_features)
    h = decoder_forward(sd, c, decoder_input_ids, enc)
    logits = F.linear(h, sd["model.decoder.embed_tokens.weight"])         # proj_out, tied, no bias (:1081)
    loss = None
    if labels is not None:
        loss = F.cross_entropy(logits.view(-1, c.vocab_size), labels.reshape(-1))   # mean, ignore_index -100
    return {"loss": loss, "logits": logits, "encoder_last_hidden_state": enc}


@torch.no_grad()
def timestamp_rules(ids, scores, begin_index, timestamp_begin, eos_token_id, max_initial_timestamp_index=None):
    """HF:generation/logits_process.py WhisperTimeStampLogitsProcessor.__call__, restated: ids int64 [B, t] (prompt + generated),
    scores fp32 [B, V] -> processed scores."""
    s = scores.clone()
    s[:, timestamp_begin - 1] = -float("inf")                                  # <|notimestamps|>
    for k in range(ids.shape[0]):
        seq = ids[k, begin_index:].tolist()
        last = len(seq) >= 1 and seq[-1] >= timestamp_begin
        penult = len(seq) < 2 or seq[-2] >= timestamp_begin
        if last:
            if penult:
                s[k, timestamp_begin:] = -float("inf")                         # a closed pair: text next
            else:
                s[k, :eos_token_id] = -float("inf")                            # text then one timestamp: no text below EOS
        ts = [t for t in seq if t >= timestamp_begin]
        if ts:
            lo = ts[-1] if (last and not penult) else ts[-1] + 1               # never decrease, never repeat <|0.00|>
            s[k, timestamp_begin:lo] = -float("inf")
    if ids.shape[1] == begin_index:
        s[:, :timestamp_begin] = -float("inf")
        if max_initial_timestamp_index is not None:
            s[:, timestamp_begin + max_initial_timestamp_index + 1:] = -float("inf")
    lp = F.log_softmax(s.float(), dim=-1)
    for k in range(ids.shape[0]):
        if lp[k, timestamp_begin:].logsumexp(dim=-1) > lp[k, :timestamp_begin].max():
            s[k, :timestamp_begin] = -float("inf")
    return s


def synthetic_timestamp_cases(B, V, ts_begin, eos, begin, seed):
    """Random logits + crafted prefixes covering every branch of the timestamp rules."""
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn((B, V), generator=g) * 2.0
    logits[:, ts_begin:] += torch.randn((B, 1), generator=g) * 2.0            # sometimes the timestamps outweigh the text
    t = lambda k: ts_begin + k                                                 # noqa: E731
    bodies = [[], [t(0)], [t(0), 7], [t(0), 7, t(2)], [t(0), 7, t(2), t(2)], [t(0), 7, t(2), t(2), 9], [t(1), 3, 4, t(3)], [t(0), t(0), 5, 6],
              [t(0), 5, t(4), t(4), 6, t(5)], [t(2), 11]]
    out = []
    for body in bodies:
        ids = torch.full((B, begin + len(body)), 5, dtype=torch.long)
        ids[:, :begin] = torch.arange(begin) + eos + 1                          # "initial tokens"
        if body:
            ids[:, begin:] = torch.tensor(body)
        out.append((ids, logits.clone()))
    return out



@torch.no_grad()
def greedy_generate(sd: dict, c: WhisperDims, input_features, prompt, eos_token_id, pad_token_id, limit, suppress_tokens=None,
                    begin_suppress_tokens=None, timestamp_begin=None, max_initial_timestamp_index=None):
    """Greedy search as HF:generation/utils.py `_sample` runs it for Whisper (do_sample False, one beam) with the two logits
    processors of HF:models/whisper/generation_whisper.py:1774-1800: `suppress_tokens` at every step, `begin_suppress_tokens` at
    the first generated position (begin_index = prompt length).  Finished rows emit pad; stops when every row has emitted EOS or
    `limit` tokens exist.  The prefix is re-decoded from scratch every step (no cache: this is the checker).
    prompt: int64 [B, P] initial tokens.  Returns int64 [B, L] including the prompt."""
    enc = encoder_forward(sd, c, input_features)
    ids = prompt.clone()
    P = prompt.shape[1]
    unfinished = torch.ones(ids.shape[0], dtype=torch.bool)
    while ids.shape[1] < limit:
        h = decoder_forward(sd, c, ids, enc)[:, -1]
        logits = F.linear(h, sd["model.decoder.embed_tokens.weight"]).float()
        if suppress_tokens:
            logits[:, list(suppress_tokens)] = -float("inf")
        if begin_suppress_tokens and ids.shape[1] == P:
            logits[:, list(begin_suppress_tokens)] = -float("inf")
        if timestamp_begin is not None:                                        # return_timestamps=True (HF :1774-1778 order)
            logits = timestamp_rules(ids, logits, P, timestamp_begin, eos_token_id, max_initial_timestamp_index)
        nxt = logits.argmax(dim=-1)
        nxt = torch.where(unfinished, nxt, torch.full_like(nxt, pad_token_id))
        ids = torch.cat([ids, nxt[:, None]], dim=1)
        unfinished &= nxt != eos_token_id
        if not bool(unfinished.any()):
            break
    return ids


def kl_divergence(target_distribution, log_predicted_distribution, labels):
    """ref:training/run_distillation.py:1453-1462."""
    divergence = F.kl_div(log_predicted_distribution, target_distribution, reduction="none")
    padding_mask = (labels >= 0).unsqueeze(-1)
    divergence = divergence * padding_mask
    return divergence.sum() / padding_mask.sum()


def kd_train_step(student_sd, student_c, teacher_sd, teacher_c, batch, temperature=2.0, kl_weight=1.0,
                  share_hidden_states=False):
    """ref:training/run_distillation.py:1465-1495.  `batch` has input_features, decoder_input_ids, labels.
    Returns (loss, metrics, student_out, teacher_out); gradients are left to the caller's autograd."""
    s_out = model_forward(student_sd, student_c, **batch)
    with torch.no_grad():
        if share_hidden_states:
            t_out = model_forward(teacher_sd, teacher_c, labels=batch["labels"],
                                  encoder_hidden_states=s_out["encoder_last_hidden_state"].detach())
        else:
            t_out = model_forward(teacher_sd, teacher_c, **batch)
    ce_loss = s_out["loss"]
    teacher_distribution = F.softmax(t_out["logits"] / temperature, dim=-1)
    student_distribution = F.log_softmax(s_out["logits"] / temperature, dim=-1)
    kl_loss = kl_divergence(teacher_distribution, student_distribution, batch["labels"]) * temperature ** 2
    loss = 0.8 * ce_loss + kl_weight * kl_loss
    return loss, {"loss": loss, "ce_loss": ce_loss, "kl_loss": kl_loss}, s_out, t_out


def collate_labels(label_rows, pad_token_id, decoder_start_token_id, max_len=None):
    """ref:training/run_distillation.py:455-476 — from padded label id rows build
    (decoder_input_ids, labels): shift, -100 on padding, -100 on the prompt up to and including SOT."""
    max_len = max_len or max(len(r) for r in label_rows)
    ids = torch.full((len(label_rows), max_len), pad_token_id, dtype=torch.long)
    att = torch.zeros((len(label_rows), max_len), dtype=torch.long)
    for i, r in enumerate(label_rows):
        ids[i, : len(r)] = torch.as_tensor(r, dtype=torch.long)
        att[i, : len(r)] = 1
    decoder_input_ids = ids[:, :-1]
    labels = ids[:, 1:]
    labels = labels.masked_fill(att[:, 1:].ne(1), -100)
    bos_index = torch.argmax((labels == decoder_start_token_id).long(), dim=1)
    bos_index = torch.where(bos_index > 0, bos_index + 1, bos_index)
    prompt_mask = torch.arange(labels.shape[1]) < bos_index[:, None]
    labels = torch.where(prompt_mask, -100, labels)
    return decoder_input_ids.contiguous(), labels.contiguous()


def synthetic_batch(c: WhisperDims, batch: int, n_tok: int, seed: int, prompt_row: bool = True):
    """Seeded synthetic KD batch: mel-like features in [-1, 1.5], random token ids with
    ragged lengths (tail -> -100) and, in row 0, a prompt prefix masked up to and including SOT so that the
    teacher's shift_tokens_right inputs differ from the student's decoder_input_ids."""
    rs = np.random.RandomState(seed)
    feats = np.clip(0.5 * rs.randn(batch, c.num_mel_bins, 2 * c.max_source_positions), -1.0, 1.5).astype(np.float32)
    hi = min(c.pad_token_id, c.decoder_start_token_id, c.vocab_size) - 1
    rows = []
    for b in range(batch):
        n = n_tok + 1 if b == batch - 1 else int(rs.randint(max(2, (n_tok + 1) // 2), n_tok + 2))
        r = rs.randint(0, hi, size=n).tolist()
        if prompt_row and b == 0 and n >= 6:
            r[3] = c.decoder_start_token_id      # tokens 0..2 act as a <|startofprev|> prompt
        else:
            r[0] = c.decoder_start_token_id
        rows.append(r)
    dec_in, labels = collate_labels(rows, c.pad_token_id, c.decoder_start_token_id, max_len=n_tok + 1)
    return {"input_features": torch.from_numpy(feats), "decoder_input_ids": dec_in, "labels": labels}
