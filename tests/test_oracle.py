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
