"""Greedy generation with a KV cache for the reference's eval loop (`student_model.generate(batch["input_features"],
**gen_kwargs)`, ref:training/run_distillation.py:1428-1446, :1524-1528) and its pseudo-labelling loop
(ref:training/run_pseudo_labelling.py:861-927: the teacher transcribes batches greedily).

What is mirrored from Hugging Face (transformers 5.5.0):
  * the initial tokens <|startoftranscript|> [language] [task] [<|notimestamps|>] from `language` / `task` / the generation
    config's lang_to_id / task_to_id / no_timestamps_token_id / forced_decoder_ids (HF:models/whisper/generation_whisper.py
    :1420-1608 `_set_language_and_task`, `_retrieve_init_tokens`), language detection when `language` is None (:1610-1674);
  * `suppress_tokens` and `begin_suppress_tokens` (HF :1774-1800 -> SuppressTokensLogitsProcessor /
    SuppressTokensAtBeginLogitsProcessor with begin_index = number of initial tokens);
  * greedy search with EOS bookkeeping (HF:generation/utils.py `_sample`: arg-max, finished rows emit pad, stop when every row
    has finished or at max_length / max_new_tokens).
The returned ids are [B, L] int64 = initial tokens + generated tokens (+ EOS, then pad), i.e. GenerationMixin.generate's
contract -- what the reference's `tokenizer.batch_decode(..., skip_special_tokens=True)` consumes.  (HF 5.x's Whisper wrapper
additionally strips the initial tokens / EOS and re-enters its long-form seek loop when a row hits max_length without EOS;
after `skip_special_tokens` the strings are the same whenever EOS is reached.)

  * `return_timestamps=True` (the reference's recommended pseudo-labelling mode, ref:training/README.md:130,148): no
    <|notimestamps|> in the initial tokens and WhisperTimeStampLogitsProcessor's rules (HF:generation/logits_process.py) inside the
    token pick (dwb_greedy_pick_timestamps), `max_initial_timestamp_index` from the generation config.
Not implemented -> NotImplementedError (never a silently different decode): beam search, sampling, prompt_ids, custom logits
processors, long-form (> 30 s) inputs.

Schedule: the encoder and the per-layer cross-attention K/V projections of its 1500 positions run once; then ONE CUDA graph
of the single-token step (embedding at position pos, per layer LN -> QKV GEMM -> cache append + attention over pos+1 keys ->
out-proj -> cross-attention over the cached encoder K/V -> MLP, LM head, token pick, advance) is replayed per token; `pos`,
the EOS flags and the token matrix live on the device, the host only polls a completion flag every few tokens.
"""
from __future__ import annotations

import torch

from . import engine, ops

BF16, F32 = torch.bfloat16, torch.float32

TASK_IDS = ("translate", "transcribe")


# ---------------------------------------------------------------------------------------------------------------------
# generation-config plumbing (host side, plain Python)
def _cfg_get(gen, name, default=None):
    if gen is None:
        return default
    if isinstance(gen, dict):
        return gen.get(name, default)
    return getattr(gen, name, default)


def _language_to_id(language: str, lang_to_id: dict) -> int:
    lang = language.lower()
    if lang in lang_to_id:
        return lang_to_id[lang]
    tok = f"<|{lang}|>"
    if tok in lang_to_id:
        return lang_to_id[tok]
    try:                                   # full names ("french") -> codes, the tokenizer's own table
        from transformers.models.whisper.tokenization_whisper import TO_LANGUAGE_CODE
        if lang in TO_LANGUAGE_CODE and f"<|{TO_LANGUAGE_CODE[lang]}|>" in lang_to_id:
            return lang_to_id[f"<|{TO_LANGUAGE_CODE[lang]}|>"]
    except Exception:  # noqa: BLE001
        pass
    raise ValueError(f"Unsupported language: {language}. It is not in `generation_config.lang_to_id`.")


def initial_tokens(model, gen, language, task, return_timestamps, detect):
    """[n_init] python ints per batch row (a list of lists when languages differ) following HF `_retrieve_init_tokens`.
    `detect()` -> LongTensor [B] of language ids, called only when a multilingual model gets language=None."""
    cfg = model.config
    multilingual = _cfg_get(gen, "is_multilingual", None)
    if multilingual is False and (task is not None or language is not None):
        raise ValueError("Cannot specify `task` or `language` for an English-only model. If the model is intended to be "
                         "multilingual, pass `is_multilingual=True` to generate, or update the generation config.")
    lang_to_id = _cfg_get(gen, "lang_to_id", None)
    task_to_id = _cfg_get(gen, "task_to_id", None)
    if language is not None and not lang_to_id:
        raise ValueError("The generation config has no `lang_to_id`: it is not compatible with the `language` argument to `generate`.")
    if task is not None and not task_to_id:
        raise ValueError("The generation config has no `task_to_id`: it is not compatible with the `task` argument to `generate`.")
    if language is None:
        language = _cfg_get(gen, "language", None)
    if task is None:
        task = _cfg_get(gen, "task", None)
    init = [_cfg_get(gen, "decoder_start_token_id", None) or cfg.decoder_start_token_id]
    if task is None and language is None:
        forced = _cfg_get(gen, "forced_decoder_ids", None) or getattr(cfg, "forced_decoder_ids", None)
        if forced:
            forced = [list(x) for x in forced]
            i = 1
            while forced and forced[0][0] == i:
                init.append(forced[0][1])
                forced = forced[1:]
                i += 1
            if forced:
                raise ValueError(f"`forced_decoder_ids` {forced} do not follow the prompt pattern of Whisper (indices must be contiguous from 1)")
    lang_undefined = len(init) <= 1 or init[1] is None
    lang_ids = None
    if language is not None:
        if isinstance(language, (list, tuple)):
            lang_ids = [_language_to_id(x, lang_to_id) for x in language]
        else:
            lang_ids = [_language_to_id(language, lang_to_id)]
    elif lang_to_id and lang_undefined:
        lang_ids = [int(x) for x in detect().tolist()]
    rows = [list(init) for _ in range(len(lang_ids) if lang_ids else 1)]
    for r, row in enumerate(rows):
        if lang_ids is not None:
            if len(row) > 1:
                row[1] = lang_ids[r]
            else:
                row.append(lang_ids[r])
        if task is not None:
            if task not in TASK_IDS:
                raise ValueError(f"The `{task}` task is not supported. The task should be one of `{TASK_IDS}`")
            tid = task_to_id[task]
            if any(t in row for t in task_to_id.values()):
                row[:] = [tid if t in task_to_id.values() else t for t in row]
            else:
                row.append(tid)
        elif language is not None and task_to_id:
            if not any(t in row for t in task_to_id.values()):
                row.append(task_to_id["transcribe"])
        nots = _cfg_get(gen, "no_timestamps_token_id", None)
        if not return_timestamps and nots is not None and row[-1] != nots:
            row.append(nots)
        elif return_timestamps and nots is not None and row[-1] == nots:
            row.pop()                    # <|notimestamps|> from forced_decoder_ids is dropped when timestamps are requested
        rows[r] = [t for t in row if t is not None]
    return rows


# ---------------------------------------------------------------------------------------------------------------------
class DecodeSession:
    """Static device state + the captured single-token graph for one (model, batch size, max length, encoder length)."""

    def __init__(self, model, B: int, t_max: int, S: int, use_graph: bool = True):
        cfg = model.config
        dec = model.model.decoder
        dev = dec.embed_tokens.weight.device
        self.model, self.B, self.t_max, self.S = model, B, t_max, S
        self.d, self.H, self.V = cfg.d_model, cfg.decoder_attention_heads, cfg.vocab_size
        d = self.d
        self.st = engine.state_of(dec)
        self.seq = torch.zeros((B, t_max), dtype=torch.int64, device=dev)
        self.pos = torch.zeros((1,), dtype=torch.int32, device=dev)
        self.finished = torch.zeros((B,), dtype=torch.int32, device=dev)
        self.done_at = torch.zeros((1,), dtype=torch.int32, device=dev)
        self.enc = torch.empty((B * S, d), dtype=BF16, device=dev)
        n_layers = len(dec.layers)
        self.cross_kv = [torch.empty((B * S, 2 * d), dtype=BF16, device=dev) for _ in range(n_layers)]
        self.self_kv = [torch.zeros((B * t_max, 2 * d), dtype=BF16, device=dev) for _ in range(n_layers)]
        self.bias_all = torch.zeros((self.V,), dtype=F32, device=dev)
        self.bias_begin = torch.zeros((self.V,), dtype=F32, device=dev)
        self.params = self._captured = None          # (prompt_len, begin_pos, eos, pad) baked into the captured pick kernel
        self.graph = None
        self.use_graph = use_graph
        self.logits = None

    # ---- one decoder step on the tokens at column pos (device scalar) --------------------------------------------
    def _step(self, prompt_len, begin_pos, eos, pad, ts_begin=None, max_initial=None):
        st, dec = self.st, self.st.m
        B, d, H, S, c = self.B, self.d, self.H, self.S, self.st.cache
        E, P = dec.embed_tokens.weight, dec.embed_positions.weight
        x = ops.embed_decode(self.seq, self.pos, E.detach(), P.detach(), d, self.V)
        y = None
        for i, layer in enumerate(dec.layers):
            k = f"l{i}"
            ws = engine._attn_weights(st, k + ".sa", layer.self_attn, fuse_qkv=True)
            g, b_ = engine._ln(st, k + ".ln1", layer.self_attn_layer_norm)
            x1, h1, _, _ = ops.add_layernorm(x, y, g, b_, rows=B, d=d)
            qkv = ops.gemm_small_m(h1, ws["wqkv"], bias=ws["bqkv"])
            kv = self.self_kv[i]
            o1 = ops.attention_decode(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], kv[:, :d], kv[:, d:], self.t_max, B, H, pos_dev=self.pos)
            y1 = ops.gemm_small_m(o1, ws["wo"], bias=ws["bo"])
            wc = engine._attn_weights(st, k + ".ca", layer.encoder_attn, fuse_qkv=False)
            g, b_ = engine._ln(st, k + ".ln2", layer.encoder_attn_layer_norm)
            x2, h2, _, _ = ops.add_layernorm(x1, y1, g, b_, rows=B, d=d)
            qc = ops.gemm_small_m(h2, wc["wq"], bias=wc["bq"])
            ckv = self.cross_kv[i]
            o2 = ops.attention_decode(qc, None, None, ckv[:, :d], ckv[:, d:], S, B, H, fixed_len=S)
            y2 = ops.gemm_small_m(o2, wc["wo"], bias=wc["bo"])
            g, b_ = engine._ln(st, k + ".ln3", layer.final_layer_norm)
            x3, h3, _, _ = ops.add_layernorm(x2, y2, g, b_, rows=B, d=d)
            a = ops.gemm_small_m(h3, engine.bf16_of(c, k + ".fc1.w", layer.fc1.weight), bias=engine.f32_of(c, k + ".fc1.b", layer.fc1.bias), act=1)
            y3 = ops.gemm_small_m(a, engine.bf16_of(c, k + ".fc2.w", layer.fc2.weight), bias=engine.f32_of(c, k + ".fc2.b", layer.fc2.bias))
            x, y = x3, y3
        g, b_ = engine._ln(st, "ln_f", dec.layer_norm)
        _, hf, _, _ = ops.add_layernorm(x, y, g, b_, rows=B, d=d, write_x=False)
        logits = engine.lm_head(st, hf)
        if ts_begin is None:
            ops.greedy_pick(logits, self.V, self.bias_all, self.bias_begin, begin_pos, self.seq, prompt_len, self.finished, eos, pad, self.pos)
        else:       # return_timestamps=True: WhisperTimeStampLogitsProcessor's rules inside the pick
            ops.greedy_pick_timestamps(logits, self.V, self.bias_all, self.bias_begin, begin_pos, self.seq, prompt_len, self.finished, eos, pad,
                                       self.pos, ts_begin, max_initial)
        ops.decode_advance(self.pos, self.finished, self.done_at)
        return logits

    def prepare(self, enc, prompt, suppress, begin_suppress, eos, pad, ts_begin=None, max_initial=None):
        """Load the encoder states, project the cross-attention K/V once per layer, reset the token matrix to the prompt."""
        st, dec = self.st, self.st.m
        d = self.d
        self.enc.copy_(enc)
        for i, layer in enumerate(dec.layers):
            wc = engine._attn_weights(st, f"l{i}.ca", layer.encoder_attn, fuse_qkv=False)
            ops.gemm(self.enc, wc["wkv"], bias=wc["bkv"], out=self.cross_kv[i])
        self._reset(prompt, pad)
        self.bias_all.zero_()
        self.bias_begin.zero_()
        if suppress:
            self.bias_all[torch.as_tensor(sorted(set(suppress)), device=self.bias_all.device)] = float("-inf")
        if begin_suppress:
            self.bias_begin[torch.as_tensor(sorted(set(begin_suppress)), device=self.bias_all.device)] = float("-inf")
        params = (prompt.shape[1], prompt.shape[1], int(eos), int(pad), ts_begin, max_initial)
        # one eager step: re-casts the bf16 shadows of weights the optimiser has changed since the last call (in place, so the
        # captured graph below keeps reading the right buffers), sets kernel attributes and warms the allocator
        self.params = params
        self._step(*params)
        self._reset(prompt, pad)
        if self.use_graph and (self.graph is None or self._captured != params):
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.logits = self._step(*params)
            self._captured = params
            self._reset(prompt, pad)

    def _reset(self, prompt, pad):
        self.seq.fill_(int(pad))
        self.seq[:, :prompt.shape[1]] = prompt
        self.pos.zero_()
        self.finished.zero_()
        self.done_at.zero_()

    def run(self, limit: int, poll_every: int = 8):
        """Decode until every row has emitted EOS or `limit` tokens exist.  Returns int64 [B, L]."""
        n_steps = limit - 1
        done = 0
        for t in range(n_steps):
            if self.graph is not None:
                self.graph.replay()
            else:
                self.logits = self._step(*self.params)
            if (t + 1) % poll_every == 0 or t + 1 == n_steps:
                done = int(self.done_at.item())           # the only host sync of the loop
                if done:
                    break
        L = min(done, limit) if done else limit
        return self.seq[:, :L].clone()


def _session(model, B, t_max, S):
    cache = model.__dict__.setdefault("_dwb_decode_sessions", {})
    key = (B, t_max, S, next(model.model.decoder.parameters()).data_ptr())
    ses = cache.get(key)
    if ses is None:
        cache.clear()                       # one live session per model: its static buffers are large
        ses = cache[key] = DecodeSession(model, B, t_max, S)
    return ses


@torch.no_grad()
def generate(model, input_features=None, decoder_input_ids=None, encoder_outputs=None, max_length=None, max_new_tokens=None,
             num_beams=None, do_sample=False, return_timestamps=None, eos_token_id=None, pad_token_id=None, language=None, task=None,
             is_multilingual=None, suppress_tokens=None, begin_suppress_tokens=None, generation_config=None, **kwargs):
    gen = generation_config if generation_config is not None else model.generation_config
    cfg = model.config
    if num_beams is None:
        num_beams = _cfg_get(gen, "num_beams", 1) or 1
    if return_timestamps is None:
        return_timestamps = bool(_cfg_get(gen, "return_timestamps", False))
    if num_beams != 1 or do_sample:
        raise NotImplementedError("only greedy decoding (num_beams=1, do_sample=False) is implemented")
    ts_begin = max_initial = None
    if return_timestamps:
        nots = _cfg_get(gen, "no_timestamps_token_id", None)
        if nots is None:
            raise ValueError("You are trying to return timestamps, but the generation config is not properly set: it has no "
                             "`no_timestamps_token_id`.")
        ts_begin = int(nots) + 1
        max_initial = _cfg_get(gen, "max_initial_timestamp_index", None)
    unsupported = [k for k in ("logits_processor", "prompt_ids", "forced_decoder_ids", "stopping_criteria", "assistant_model",
                               "prefix_allowed_tokens_fn", "temperature", "attention_mask") if kwargs.get(k) is not None]
    if unsupported:
        raise NotImplementedError(f"generate(): unsupported arguments {unsupported}")
    unknown = [k for k in kwargs if kwargs[k] is not None and k not in ("use_cache", "return_dict_in_generate", "synced_gpus")]
    if unknown:
        raise NotImplementedError(f"generate(): unknown arguments {unknown} (refusing to ignore them)")
    if kwargs.get("return_dict_in_generate"):
        raise NotImplementedError("generate(): return_dict_in_generate is not implemented")
    if is_multilingual is not None:
        gen = dict(_gen_as_dict(gen), is_multilingual=bool(is_multilingual))
    was_training = model.training
    model.eval()
    try:
        enc_in = None
        if encoder_outputs is not None:
            enc_in = encoder_outputs[0] if isinstance(encoder_outputs, (tuple, list)) else encoder_outputs.last_hidden_state
        elif input_features is not None and input_features.shape[-1] > 2 * cfg.max_source_positions:
            raise NotImplementedError("long-form generation (> 30 s of features) is not implemented")
        enc, S, _ = engine.run_encoder(model, input_features, enc_in)
        B = enc.shape[0] // S
        dev = enc.device
        st = engine.state_of(model.model.decoder)

        def detect():      # HF :1610-1674: one decoder step on <|startoftranscript|>, arg-max over the language tokens
            sot = torch.full((B, 1), cfg.decoder_start_token_id, dtype=torch.long, device=dev)
            hf, _ = engine.decoder_forward(st, sot, enc, B, S, save=False)
            logits = engine.lm_head(st, hf)[:, :cfg.vocab_size]
            ids = torch.as_tensor(sorted(set(_cfg_get(gen, "lang_to_id").values())), device=dev)
            return ids[logits[:, ids].argmax(dim=-1)]

        if decoder_input_ids is not None:
            prompt = decoder_input_ids.to(device=dev, dtype=torch.long)
        else:
            rows = initial_tokens(model, gen, language, task, return_timestamps, detect)
            if len(rows) not in (1, B):
                raise ValueError(f"When passing a list of languages, its length must match the batch size ({B}), got {len(rows)}")
            prompt = torch.as_tensor(rows, dtype=torch.long, device=dev).expand(B, -1)
        prompt = prompt.contiguous()
        P = prompt.shape[1]
        eos = eos_token_id if eos_token_id is not None else _cfg_get(gen, "eos_token_id", None)
        if eos is None:
            eos = cfg.eos_token_id
        if isinstance(eos, (list, tuple)):
            if len(eos) != 1:
                raise NotImplementedError("several EOS token ids are not implemented")
            eos = eos[0]
        pad = pad_token_id if pad_token_id is not None else _cfg_get(gen, "pad_token_id", None)
        if pad is None:
            pad = cfg.pad_token_id
        if max_new_tokens is None:
            max_new_tokens = _cfg_get(gen, "max_new_tokens", None) if max_length is None else None
        if max_new_tokens is not None:
            limit = P + int(max_new_tokens)
        else:
            limit = int(max_length or _cfg_get(gen, "max_length", None) or cfg.max_length)
        limit = min(limit, cfg.max_target_positions)
        if limit <= P:
            return prompt
        if suppress_tokens is None:
            suppress_tokens = _cfg_get(gen, "suppress_tokens", None)
        if begin_suppress_tokens is None:
            begin_suppress_tokens = _cfg_get(gen, "begin_suppress_tokens", None)
        ses = _session(model, B, limit, S)
        ses.prepare(enc, prompt, suppress_tokens, begin_suppress_tokens, eos, pad, ts_begin, max_initial)
        return ses.run(limit)
    finally:
        model.train(was_training)


def _gen_as_dict(gen):
    if gen is None:
        return {}
    if isinstance(gen, dict):
        return dict(gen)
    if hasattr(gen, "to_dict"):
        return gen.to_dict()
    return dict(vars(gen))
