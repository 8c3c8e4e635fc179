"""Pin oracle/*.py to the golden vectors minted from the installed HF 5.5.0 implementation
(oracle/gen_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import logmel_oracle as lo
from oracle import whisper_oracle as wo


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


@pytest.mark.parametrize("branch,share", [("A", False), ("B", True)])
def test_kd_step_matches_hf_golden(golden_dir, branch, share):
    g = _load(golden_dir, "kd_tiny.npz")
    s_seed, t_seed, b_seed = [int(x) for x in g["seeds"]]
    sc, tc = wo.PRESETS["tiny-student"], wo.PRESETS["tiny-teacher"]
    ssd = wo.init_state_dict(sc, s_seed)
    tsd = wo.init_state_dict(tc, t_seed)
    if share:
        for k in list(tsd):
            if k.startswith("model.encoder."):
                tsd[k] = ssd[k]
    for k, v in ssd.items():
        v.requires_grad_(not (share and k.startswith("model.encoder.")) and k != "model.encoder.embed_positions.weight")
    batch = wo.synthetic_batch(sc, batch=3, n_tok=12, seed=b_seed)
    # the prompt-masked row makes teacher decoder inputs differ from the student's (SURVEY.md 8a note)
    assert (batch["labels"][0, :3] == -100).all() and batch["labels"][0, 3] >= 0
    loss, metrics, so, to = wo.kd_train_step(ssd, sc, tsd, tc, batch, 2.0, 1.0, share_hidden_states=share)
    loss.backward()
    np.testing.assert_allclose(loss.item(), g[f"{branch}_loss"], rtol=2e-6)
    np.testing.assert_allclose(metrics["ce_loss"].item(), g[f"{branch}_ce_loss"], rtol=2e-6)
    np.testing.assert_allclose(metrics["kl_loss"].item(), g[f"{branch}_kl_loss"], rtol=2e-5)
    np.testing.assert_allclose(so["logits"].detach().numpy(), g[f"{branch}_student_logits"], atol=2e-6, rtol=1e-4)
    np.testing.assert_allclose(to["logits"].detach().numpy(), g[f"{branch}_teacher_logits"], atol=2e-6, rtol=1e-4)
    np.testing.assert_allclose(so["encoder_last_hidden_state"].detach().numpy(),
                               g[f"{branch}_encoder_last_hidden_state"], atol=1e-5, rtol=1e-4)
    assert (so["logits"].argmax(-1).numpy() == g[f"{branch}_student_logits"].argmax(-1)).all()
    names = [str(n) for n in g[f"{branch}_grad_names"]]
    got = {k for k, v in ssd.items() if v.grad is not None}
    assert got == set(names)
    for n, norm, sm in zip(names, g[f"{branch}_grad_norms"], g[f"{branch}_grad_sums"]):
        gr = ssd[n].grad
        np.testing.assert_allclose(float(gr.norm()), norm, rtol=2e-4, atol=1e-9, err_msg=n)
        key = f"{branch}_grad::{n}"
        if key in g.files:
            np.testing.assert_allclose(gr.numpy(), g[key], atol=2e-6 * max(1.0, float(norm)), rtol=1e-3, err_msg=n)
    # padding row of the embedding only gets gradient through the tied LM head (SURVEY.md "hard parts")
    assert float(ssd["model.decoder.embed_tokens.weight"].grad[sc.pad_token_id].abs().sum()) > 0


def test_shift_and_collate_semantics():
    c = wo.PRESETS["tiny-student"]
    rows = [[7, 8, 9, c.decoder_start_token_id, 1, 2, 3], [c.decoder_start_token_id, 4, 5]]
    dec_in, labels = wo.collate_labels(rows, c.pad_token_id, c.decoder_start_token_id)
    assert dec_in.shape == labels.shape == (2, 6)
    assert labels[0].tolist() == [-100, -100, -100, 1, 2, 3]
    assert labels[1].tolist() == [4, 5, -100, -100, -100, -100]
    assert dec_in[1].tolist() == [c.decoder_start_token_id, 4, 5, c.pad_token_id, c.pad_token_id, c.pad_token_id]
    sh = wo.shift_tokens_right(labels, c.pad_token_id, c.decoder_start_token_id)
    assert sh[0].tolist() == [c.decoder_start_token_id, c.pad_token_id, c.pad_token_id, c.pad_token_id, 1, 2]
    assert sh[1].tolist() == dec_in[1].tolist()


def test_encoder_length_check():
    c = wo.PRESETS["tiny-student"]
    sd = wo.init_state_dict(c, 1)
    with pytest.raises(ValueError):
        wo.encoder_forward(sd, c, torch.zeros(1, c.num_mel_bins, 2 * c.max_source_positions - 2))


@pytest.mark.parametrize("n_mels", [80, 128])
def test_logmel_matches_hf_golden(golden_dir, n_mels):
    g = _load(golden_dir, f"logmel_{n_mels}.npz")
    np.testing.assert_allclose(lo.mel_filter_bank(n_mels), g["mel_filters"], atol=1e-12)
    wav = lo.synthetic_waveforms(3, seed=int(g["seed"]), ragged=False)
    wav[1, 200000:] = 0.0
    clips = lo.pad_or_trim([wav[0], wav[1], wav[2, : int(g["short_len"])]])
    out = lo.log_mel(clips, n_mels=n_mels)
    assert out.shape == (3, n_mels, 3000) and out.dtype == np.float32
    # HF documents 1e-5 between its own numpy and torch STFT paths (HF:feature_extraction_whisper.py:107-108)
    np.testing.assert_allclose(out[:, :, g["frames"]], g["values"], atol=2e-5)
    np.testing.assert_allclose(out.mean(axis=2), g["row_mean"], atol=2e-6)
    np.testing.assert_allclose(out.reshape(3, -1).max(axis=1), g["utt_max"], atol=2e-5)
    np.testing.assert_allclose(out.reshape(3, -1).min(axis=1), g["utt_min"], atol=2e-5)


# ---- greedy generation (eval / pseudo-labelling decode) -----------------------------------------------------------------
GEN_MULTI = dict(decoder_start_token_id=501, eos_token_id=502, pad_token_id=500, bos_token_id=502,
                 suppress_tokens=[1, 2, 7, 8, 9, 220], begin_suppress_tokens=[220, 502], is_multilingual=True,
                 lang_to_id={"<|en|>": 503, "<|fr|>": 504, "<|de|>": 505}, task_to_id={"transcribe": 506, "translate": 507},
                 no_timestamps_token_id=508)
GEN_EN = dict(decoder_start_token_id=501, eos_token_id=502, pad_token_id=500, bos_token_id=502, is_multilingual=False,
              no_timestamps_token_id=508, suppress_tokens=None, begin_suppress_tokens=None)
GEN_CASES = {"A": (GEN_MULTI, dict(language="fr", task="transcribe", max_new_tokens=10)),
             "B": (GEN_MULTI, dict(max_length=12)),
             "C": (GEN_EN, dict(max_new_tokens=6)),
             "D": (dict(GEN_MULTI, max_initial_timestamp_index=1), dict(language="en", task="transcribe", max_new_tokens=8, return_timestamps=True))}


def test_oracle_greedy_generate_reproduces_hf_golden(golden_dir):
    """HF's own initial-token logic + suppress processors + greedy search (oracle/gen_golden.py:hf_greedy) on the tiny model:
    the oracle restatement must emit the same token ids, row for row, including EOS / pad bookkeeping and early stop."""
    import torch
    from oracle import whisper_oracle as wo
    g = np.load(os.path.join(golden_dir, "generate_tiny.npz"))
    sc = wo.PRESETS["tiny-student"]
    sd = wo.init_state_dict(sc, int(g["model_seed"]), std=float(g["model_std"]))
    for name, (cfg, kw) in GEN_CASES.items():
        feats, init, seq = torch.from_numpy(g[f"{name}_feats"]), torch.from_numpy(g[f"{name}_init"]), g[f"{name}_seq"]
        limit = kw["max_length"] if "max_length" in kw else init.shape[1] + kw["max_new_tokens"]
        ts = dict(timestamp_begin=cfg["no_timestamps_token_id"] + 1, max_initial_timestamp_index=cfg.get("max_initial_timestamp_index")) \
            if kw.get("return_timestamps") else {}
        out = wo.greedy_generate(sd, sc, feats, init, int(g[f"{name}_eos"]), cfg["pad_token_id"], limit, cfg.get("suppress_tokens"),
                                 [int(t) for t in g[f"{name}_begin_suppress"]], **ts)
        assert out.shape == seq.shape and (out.numpy() == seq).all(), (name, out, seq)
        assert float(g[f"{name}_min_rel_margin"].min()) > 0.02       # the fixture rows were chosen for decisive arg-maxima


def test_initial_tokens_follow_hf_retrieve_init_tokens(golden_dir):
    """Host logic of distil_whisper_b200.generation.initial_tokens == HF `_retrieve_init_tokens` (golden `*_init`) for: forced
    language + task, language detection without a task, an English-only config; plus the error cases HF raises on."""
    import types

    import pytest
    import torch
    from distil_whisper_b200 import generation
    g = np.load(os.path.join(golden_dir, "generate_tiny.npz"))
    model = types.SimpleNamespace(config=types.SimpleNamespace(decoder_start_token_id=501, forced_decoder_ids=None))
    a = generation.initial_tokens(model, GEN_MULTI, "fr", "transcribe", False, detect=None)
    assert a == [g["A_init"][0].tolist()]
    assert generation.initial_tokens(model, GEN_MULTI, "french", "translate", False, None) == [[501, 504, 507, 508]]
    assert generation.initial_tokens(model, GEN_MULTI, "<|de|>", None, False, None) == [[501, 505, 506, 508]]
    assert generation.initial_tokens(model, GEN_MULTI, ["en", "de"], None, True, None) == [[501, 503, 506], [501, 505, 506]]
    det = torch.as_tensor(g["B_init"][:, 1])
    b = generation.initial_tokens(model, GEN_MULTI, None, None, False, detect=lambda: det)
    assert b == g["B_init"].tolist()
    assert generation.initial_tokens(model, GEN_EN, None, None, False, None) == [g["C_init"][0].tolist()]
    assert generation.initial_tokens(model, GEN_MULTI, "en", "transcribe", True, None) == [g["D_init"][0].tolist()]     # no <|notimestamps|>
    forced_nots = dict(GEN_MULTI, forced_decoder_ids=[[1, 504], [2, 507], [3, 508]])
    assert generation.initial_tokens(model, forced_nots, None, None, True, None) == [[501, 504, 507]]
    forced = dict(GEN_MULTI, forced_decoder_ids=[[1, 504], [2, 507], [3, 508]])
    assert generation.initial_tokens(model, forced, None, None, False, None) == [[501, 504, 507, 508]]
    with pytest.raises(ValueError):
        generation.initial_tokens(model, GEN_EN, "en", None, False, None)          # English-only model + language
    with pytest.raises(ValueError):
        generation.initial_tokens(model, GEN_MULTI, "xx", None, False, None)       # unknown language
    with pytest.raises(ValueError):
        generation.initial_tokens(model, GEN_MULTI, "en", "summarise", False, None)


def test_oracle_timestamp_rules_match_hf_processor():
    """oracle.timestamp_rules == transformers' WhisperTimeStampLogitsProcessor on crafted prefixes (pins the restatement the
    GPU pick kernel is checked against)."""
    import types

    import pytest
    import torch
    transformers = pytest.importorskip("transformers")
    from transformers.generation import WhisperTimeStampLogitsProcessor
    from oracle import whisper_oracle as wo
    V, ts_begin, eos, begin = 140, 130, 120, 3
    for max_init in (None, 1):
        cfg = types.SimpleNamespace(no_timestamps_token_id=ts_begin - 1, eos_token_id=eos, bos_token_id=eos, max_initial_timestamp_index=max_init)
        proc = WhisperTimeStampLogitsProcessor(cfg, begin_index=begin)
        for ids, logits in wo.synthetic_timestamp_cases(6, V, ts_begin, eos, begin, seed=3):
            want = proc(ids, logits.clone())
            got = wo.timestamp_rules(ids, logits.clone(), begin, ts_begin, eos, max_init)
            assert torch.equal(torch.isinf(want), torch.isinf(got)) and torch.equal(want.argmax(-1), got.argmax(-1))
