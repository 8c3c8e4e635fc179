"""Generate tests/golden/*.npz from the INSTALLED Hugging Face implementation (transformers 5.5.0).

Run in the build container only:  python oracle/gen_golden.py
The reference repo has no tests, fixtures or golden vectors for this path; the
arithmetic lives in `transformers` (ref:training/setup.py:22 pins >=4.35.1).  This script therefore drives
the HF classes with a literal copy of ref:training/run_distillation.py:1453-1495 and records what they
produce on seeded tiny inputs; tests/test_oracle.py pins oracle/*.py to these files and the -m gpu tests
compare the CUDA path to the oracle (and to these files directly).

Weights are NOT stored: both sides rebuild them from numpy's frozen RandomState via
oracle.whisper_oracle.init_state_dict(dims, seed).
"""
import os
import sys

import numpy as np
import torch
from torch import nn

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import whisper_oracle as wo      # noqa: E402
from oracle import logmel_oracle as lo       # noqa: E402

OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
STUDENT_SEED, TEACHER_SEED, BATCH_SEED = 11, 23, 5


def hf_model(dims: wo.WhisperDims, sd: dict):
    import transformers
    from transformers import WhisperConfig, WhisperForConditionalGeneration
    cfg = WhisperConfig(
        vocab_size=dims.vocab_size, num_mel_bins=dims.num_mel_bins, d_model=dims.d_model,
        encoder_layers=dims.encoder_layers, encoder_attention_heads=dims.encoder_attention_heads,
        encoder_ffn_dim=dims.encoder_ffn_dim, decoder_layers=dims.decoder_layers,
        decoder_attention_heads=dims.decoder_attention_heads, decoder_ffn_dim=dims.decoder_ffn_dim,
        max_source_positions=dims.max_source_positions, max_target_positions=dims.max_target_positions,
        pad_token_id=dims.pad_token_id, bos_token_id=dims.pad_token_id, eos_token_id=dims.pad_token_id,
        decoder_start_token_id=dims.decoder_start_token_id, suppress_tokens=None, begin_suppress_tokens=None,
    )
    m = WhisperForConditionalGeneration(cfg)
    full = dict(sd)
    full["proj_out.weight"] = sd["model.decoder.embed_tokens.weight"]
    missing, unexpected = m.load_state_dict(full, strict=False)
    assert not unexpected, unexpected
    assert all("proj_out" in k for k in missing), missing
    assert m.proj_out.weight.data_ptr() == m.model.decoder.embed_tokens.weight.data_ptr(), "tie lost"
    return m, transformers.__version__


def reference_train_step(student_model, teacher_model, batch, temperature, kl_weight, share_hidden_states):
    """Literal restatement of ref:training/run_distillation.py:1453-1495 on HF modules."""
    from transformers.modeling_outputs import BaseModelOutput

    def kl_divergence(target_distribution, log_predicted_distribution, labels):
        kl_loss = nn.KLDivLoss(reduction="none")
        divergence = kl_loss(log_predicted_distribution, target_distribution)
        padding_mask = labels >= 0
        padding_mask = padding_mask.unsqueeze(-1)
        divergence = divergence * padding_mask
        divergence = divergence.sum() / padding_mask.sum()
        return divergence

    student_model.train()
    teacher_model.eval()
    student_outputs = student_model(**batch)
    with torch.no_grad():
        if share_hidden_states:
            encoder_outputs = BaseModelOutput(student_outputs.encoder_last_hidden_state)
            teacher_outputs = teacher_model(encoder_outputs=encoder_outputs, labels=batch["labels"])
        else:
            teacher_outputs = teacher_model(**batch)
    ce_loss = student_outputs.loss
    teacher_distribution = nn.functional.softmax(teacher_outputs.logits / temperature, dim=-1)
    student_distribution = nn.functional.log_softmax(student_outputs.logits / temperature, dim=-1)
    kl_loss = kl_divergence(teacher_distribution, student_distribution, batch["labels"]) * temperature**2
    loss = 0.8 * ce_loss + kl_weight * kl_loss
    return loss, {"loss": loss, "ce_loss": ce_loss, "kl_loss": kl_loss}, student_outputs, teacher_outputs


def gen_kd():
    torch.set_num_threads(4)
    sc, tc = wo.PRESETS["tiny-student"], wo.PRESETS["tiny-teacher"]
    out = {}
    for branch, share in (("A", False), ("B", True)):
        ssd = wo.init_state_dict(sc, STUDENT_SEED)
        tsd = wo.init_state_dict(tc, TEACHER_SEED)
        if share:   # ref:training/run_distillation.py:1046-1049: teacher.model.encoder = student.model.encoder
            for k in list(tsd):
                if k.startswith("model.encoder."):
                    tsd[k] = ssd[k]
        student, ver = hf_model(sc, ssd)
        teacher, _ = hf_model(tc, tsd)
        if share:
            for p in student.model.encoder.parameters():
                p.requires_grad = False
        batch = wo.synthetic_batch(sc, batch=3, n_tok=12, seed=BATCH_SEED)
        loss, metrics, so, to = reference_train_step(student, teacher, batch, 2.0, 1.0, share)
        loss.backward()
        out[f"{branch}_loss"] = loss.detach().numpy()
        out[f"{branch}_ce_loss"] = metrics["ce_loss"].detach().numpy()
        out[f"{branch}_kl_loss"] = metrics["kl_loss"].detach().numpy()
        out[f"{branch}_student_logits"] = so.logits.detach().numpy()
        out[f"{branch}_teacher_logits"] = to.logits.detach().numpy()
        out[f"{branch}_encoder_last_hidden_state"] = so.encoder_last_hidden_state.detach().numpy()
        names, norms, sums = [], [], []
        for n, p in student.named_parameters():
            if p.grad is None:
                continue
            names.append(n)
            norms.append(float(p.grad.norm()))
            sums.append(float(p.grad.double().sum()))
            keep = (p.grad.numel() <= 17000 and (".layers." not in n or ".layers.0." in n)) or n.endswith("embed_tokens.weight")
            if keep:
                out[f"{branch}_grad::{n}"] = p.grad.detach().numpy()
        out[f"{branch}_grad_names"] = np.array(names)
        out[f"{branch}_grad_norms"] = np.array(norms, dtype=np.float64)
        out[f"{branch}_grad_sums"] = np.array(sums, dtype=np.float64)
        print(branch, "loss", float(loss), "ce", float(metrics["ce_loss"]), "kl", float(metrics["kl_loss"]),
              "n_grads", len(names))
    out["transformers_version"] = np.array(ver)
    out["torch_version"] = np.array(torch.__version__)
    out["seeds"] = np.array([STUDENT_SEED, TEACHER_SEED, BATCH_SEED])
    np.savez_compressed(os.path.join(OUT, "kd_tiny.npz"), **out)


def gen_logmel():
    from transformers import WhisperFeatureExtractor
    import transformers
    wav = lo.synthetic_waveforms(3, seed=7, ragged=False)
    wav[1, 200000:] = 0.0                       # hard-zero tail -> clamp + floor path
    short = wav[2, :123457].copy()              # short clip -> extractor zero-pads (HF:...:296)
    for n_mels in (80, 128):
        fe = WhisperFeatureExtractor(feature_size=n_mels)
        feats = fe([wav[0], wav[1], short], sampling_rate=16000, return_tensors="np")["input_features"]
        assert feats.shape == (3, n_mels, 3000), feats.shape
        sel = np.r_[0:48, 1250:1298, 1476:1500 + 24, 2952:3000]
        np.savez_compressed(
            os.path.join(OUT, f"logmel_{n_mels}.npz"),
            frames=sel, values=feats[:, :, sel].astype(np.float32),
            row_mean=feats.mean(axis=2).astype(np.float64), utt_max=feats.reshape(3, -1).max(axis=1),
            utt_min=feats.reshape(3, -1).min(axis=1),
            mel_filters=fe.mel_filters.astype(np.float64),
            transformers_version=np.array(transformers.__version__), seed=np.array(7), short_len=np.array(123457))
        print("logmel", n_mels, feats.mean(), feats.min(), feats.max())


# ---------------------------------------------------------------------------------------------------------------------
# greedy generation (the eval / pseudo-labelling decode): HF's own initial-token logic, logits processors and greedy search
GEN_MODEL_SEED, GEN_MODEL_STD, GEN_POOL_SEED, GEN_POOL = 4, 0.3, 99, 96
GEN_MULTI = dict(decoder_start_token_id=501, eos_token_id=502, pad_token_id=500, bos_token_id=502,
                 suppress_tokens=[1, 2, 7, 8, 9, 220], begin_suppress_tokens=[220, 502], is_multilingual=True,
                 lang_to_id={"<|en|>": 503, "<|fr|>": 504, "<|de|>": 505}, task_to_id={"transcribe": 506, "translate": 507},
                 no_timestamps_token_id=508)
GEN_EN = dict(decoder_start_token_id=501, eos_token_id=502, pad_token_id=500, bos_token_id=502, is_multilingual=False,
              no_timestamps_token_id=508, suppress_tokens=None, begin_suppress_tokens=None)


def hf_greedy(m, feats, gen_cfg: dict, language, task, eos, timestamps=False, **length):
    """Initial tokens from WhisperGenerationMixin's own helpers (HF:models/whisper/generation_whisper.py:1420-1608), then
    GenerationMixin.generate (greedy `_sample`) with HF's suppress processors -- the 4.3x-era contract the reference was written
    against: prompt + generated (+ EOS, pad).  HF 5.x's Whisper wrapper post-processes this (strips prompt / EOS, long-form seek)."""
    from transformers import GenerationConfig
    from transformers.generation import (GenerationMixin, LogitsProcessorList, SuppressTokensAtBeginLogitsProcessor,
                                         SuppressTokensLogitsProcessor)
    g = GenerationConfig(**{k: v for k, v in dict(gen_cfg, eos_token_id=eos).items() if v is not None})
    g.return_timestamps = bool(timestamps)
    m._set_language_and_task(language=language, task=task, is_multilingual=None, generation_config=g)
    init = m._retrieve_init_tokens(feats, batch_size=feats.shape[0], generation_config=g, config=m.config,
                                   num_segment_frames=feats.shape[-1], kwargs={})
    procs = []
    if gen_cfg.get("suppress_tokens"):
        procs.append(SuppressTokensLogitsProcessor(gen_cfg["suppress_tokens"], device="cpu"))
    begin = [eos if t == gen_cfg["eos_token_id"] else t for t in (gen_cfg.get("begin_suppress_tokens") or [])]
    if begin:
        procs.append(SuppressTokensAtBeginLogitsProcessor(begin, begin_index=init.shape[1], device="cpu"))
    if timestamps:       # HF :1774-1778: the timestamp processor runs after the two suppress processors
        from transformers.generation import WhisperTimeStampLogitsProcessor
        procs.append(WhisperTimeStampLogitsProcessor(g, begin_index=init.shape[1]))
    g2 = GenerationConfig(decoder_start_token_id=gen_cfg["decoder_start_token_id"], eos_token_id=eos, pad_token_id=gen_cfg["pad_token_id"],
                          bos_token_id=gen_cfg["bos_token_id"], do_sample=False, num_beams=1, **length)
    plist = LogitsProcessorList(procs)
    out = GenerationMixin.generate(m, feats, decoder_input_ids=init, logits_processor=plist, generation_config=g2)
    hf_greedy.last_processors = plist
    return init, out, begin


def _row_margins(m, feats, out, P, suppress, begin):
    """Per row: min over generated positions of (top1 - top2) of the processed fp32 logits along the decoded path, relative to
    the largest |logit| -- rows are independent under greedy search, so the fixture keeps the rows with the safest margins."""
    with torch.no_grad():
        lg = m(input_features=feats, decoder_input_ids=out[:, :-1]).logits.float()
    if suppress:
        lg[:, :, suppress] = -1e9
    if begin:
        lg[:, P - 1, begin] = -1e9
    mx = float(lg[lg > -1e8].abs().max())
    t2 = lg[:, P - 1:].topk(2, -1).values
    return (t2[..., 0] - t2[..., 1]).min(dim=1).values / mx, lg


def gen_generate(n_rows=4):
    sc = wo.PRESETS["tiny-student"]
    sd = wo.init_state_dict(sc, GEN_MODEL_SEED, std=GEN_MODEL_STD)
    m, ver = hf_model(sc, sd)
    m.eval()
    g = torch.Generator().manual_seed(GEN_POOL_SEED)
    pool = (0.5 * torch.randn((GEN_POOL, sc.num_mel_bins, 2 * sc.max_source_positions), generator=g)).clamp_(-1.0, 1.5)
    out = {"transformers_version": np.array(ver), "model_seed": np.array(GEN_MODEL_SEED), "model_std": np.array(GEN_MODEL_STD)}
    NEVER = 509

    def case(name, gen_cfg, language, task, pick_eos, timestamps=False, **length):
        if timestamps:      # margins of the PROCESSED scores (the rules are path dependent) along each row's decoded path; golden EOS
            eos = gen_cfg["eos_token_id"]
            init, seq, begin = hf_greedy(m, pool, gen_cfg, language, task, eos, timestamps=True, **length)
            procs, P = hf_greedy.last_processors, init.shape[1]
            with torch.no_grad():
                lg = m(input_features=pool, decoder_input_ids=seq[:, :-1]).logits.float()
            mx = float(lg.abs().max())
            marg = torch.full((GEN_POOL,), 1e9)
            for t in range(P, seq.shape[1]):
                sc = procs(seq[:, :t], lg[:, t - 1].clone())
                t2 = sc.topk(2, -1).values
                gap = (t2[:, 0] - t2[:, 1]) / mx
                alive = ~(seq[:, P:t] == eos).any(dim=1)                      # positions after EOS are padding
                marg = torch.where(alive, torch.minimum(marg, gap), marg)
            rows = marg.topk(n_rows).indices.sort().values
            feats = pool[rows]
            init2, seq2, begin2 = hf_greedy(m, feats, gen_cfg, language, task, eos, timestamps=True, **length)
            for r in range(n_rows):
                a_, b_ = seq2[r].tolist(), seq[rows[r]].tolist()
                stop = a_.index(eos) + 1 if eos in a_[P:] else len(a_)
                assert a_[:stop] == b_[:stop], (name, r, a_, b_)
            out[f"{name}_feats"] = feats.numpy().astype(np.float32)
            out[f"{name}_init"] = init2.numpy()
            out[f"{name}_seq"] = seq2.numpy()
            out[f"{name}_eos"] = np.array(eos)
            out[f"{name}_begin_suppress"] = np.array(begin2, dtype=np.int64)
            out[f"{name}_min_rel_margin"] = marg[rows].numpy()
            print("generate", name, "(timestamps) rows", rows.tolist(), "min rel margin", marg[rows].min().item(), "\n", seq2)
            return
        init, seq, begin = hf_greedy(m, pool, gen_cfg, language, task, NEVER, **length)
        P = init.shape[1]
        marg, lg = _row_margins(m, pool, seq, P, gen_cfg.get("suppress_tokens"), [NEVER if t == gen_cfg["eos_token_id"] else t for t in (gen_cfg.get("begin_suppress_tokens") or [])])
        if language is None and gen_cfg.get("lang_to_id"):        # language detection must be decisive too
            with torch.no_grad():
                l0 = m(input_features=pool, decoder_input_ids=torch.full((GEN_POOL, 1), 501)).logits[:, 0, sorted(gen_cfg["lang_to_id"].values())]
            t2 = l0.topk(2, -1).values
            marg = torch.minimum(marg, (t2[:, 0] - t2[:, 1]) / float(lg[lg > -1e8].abs().max()))
        rows = marg.topk(n_rows).indices.sort().values
        feats = pool[rows]
        eos = NEVER
        if pick_eos:      # declare a token that one kept row emits mid-sequence to be EOS: rows then stop at different lengths
            eos = int(seq[rows[0], P + 3])
        init2, seq2, begin2 = hf_greedy(m, feats, gen_cfg, language, task, eos, **length)
        # rows are independent: the same rows decoded alone must reproduce their pool decode up to the first EOS
        for r in range(n_rows):
            a, b = seq2[r].tolist(), seq[rows[r]].tolist()
            stop = a.index(eos) + 1 if eos in a else len(a)
            assert a[:stop] == b[:stop], (name, r, a, b)
        out[f"{name}_feats"] = feats.numpy().astype(np.float32)
        out[f"{name}_init"] = init2.numpy()
        out[f"{name}_seq"] = seq2.numpy()
        out[f"{name}_eos"] = np.array(eos)
        out[f"{name}_begin_suppress"] = np.array(begin2, dtype=np.int64)
        out[f"{name}_min_rel_margin"] = marg[rows].numpy()
        print("generate", name, "rows", rows.tolist(), "min rel margin", marg[rows].min().item(), "eos", eos, "\n", seq2)

    case("A", GEN_MULTI, "fr", "transcribe", True, max_new_tokens=10)          # ref gen_kwargs for multilingual models (:1441-1445)
    case("B", GEN_MULTI, None, None, False, max_length=12)                     # language detection, no task, max_length semantics
    case("C", GEN_EN, None, None, True, max_new_tokens=6)                      # English-only model: <|sot|><|notimestamps|>
    case("D", dict(GEN_MULTI, max_initial_timestamp_index=1), "en", "transcribe", False, timestamps=True, max_new_tokens=8)   # pseudo-labelling mode
    np.savez_compressed(os.path.join(OUT, "generate_tiny.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["kd", "logmel", "generate"]
    if "kd" in which:
        gen_kd()
    if "logmel" in which:
        gen_logmel()
    if "generate" in which:
        gen_generate()
