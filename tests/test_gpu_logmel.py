"""Log-mel kernel (cluster + DSMEM + smem FFT) against the numpy oracle and the committed HF golden values."""
import os

import numpy as np
import pytest
import torch

from oracle import logmel_oracle as lo

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

# fp32 mixed-radix FFT vs the float64 oracle: |d log10| <= ~1e-4 for bins 80 dB below the peak, /4 in the output.
# HF states 1e-5 between its own two float32 paths on speech (HF:feature_extraction_whisper.py:107-108).
ATOL = 5e-5


@pytest.fixture(scope="module")
def fe():
    from distil_whisper_b200.feature_extraction import WhisperFeatureExtractorB200
    return {80: WhisperFeatureExtractorB200(80), 128: WhisperFeatureExtractorB200(128)}


@pytest.mark.parametrize("n_mels", [80, 128])
def test_logmel_matches_hf_golden(fe, golden_dir, n_mels):
    g = np.load(os.path.join(golden_dir, f"logmel_{n_mels}.npz"))
    np.testing.assert_allclose(fe[n_mels].mel_filters, g["mel_filters"], atol=1e-12)
    wav = lo.synthetic_waveforms(3, seed=int(g["seed"]), ragged=False)
    wav[1, 200000:] = 0.0
    out = fe[n_mels]([wav[0], wav[1], wav[2, : int(g["short_len"])]], sampling_rate=16000, return_tensors="np")["input_features"]
    assert out.shape == (3, n_mels, 3000) and out.dtype == np.float32
    np.testing.assert_allclose(out[:, :, g["frames"]], g["values"], atol=ATOL)
    np.testing.assert_allclose(out.mean(axis=2), g["row_mean"], atol=1e-5)
    np.testing.assert_allclose(out.reshape(3, -1).max(axis=1), g["utt_max"], atol=ATOL)
    np.testing.assert_allclose(out.reshape(3, -1).min(axis=1), g["utt_min"], atol=ATOL)


def test_logmel_matches_oracle_full(fe):
    wav = lo.synthetic_waveforms(6, seed=3, ragged=True)
    wav[0] *= 10.0
    wav[1] = 0.0                                   # silent clip: everything sits on the 1e-10 clamp
    wav[2, :1000] = 0.0
    t = np.arange(lo.N_SAMPLES) / 16000.0
    wav[3] = (0.3 * np.sin(2 * np.pi * 440.0 * t) + 0.01 * np.sin(2 * np.pi * 3000.0 * t)).astype(np.float32)   # tonal
    ref = lo.log_mel(wav, 80)
    out = fe[80].extract_device(torch.from_numpy(wav).cuda()).cpu().numpy()
    err = np.abs(out - ref)
    # tonal clip: leakage bins ~100 dB under the peak carry fp32 FFT noise; they are floored by max-8 anyway
    assert err.max() < 2e-4, (err.max(), np.unravel_index(err.argmax(), err.shape))
    assert np.quantile(err, 0.999) < ATOL
    assert np.all(out[1] == out[1, 0, 0])          # silence -> constant (log10(1e-10)+4)/4 = -1.5
    np.testing.assert_allclose(out[1, 0, 0], -1.5, atol=1e-6)


def test_logmel_rejects_bad_rate(fe):
    with pytest.raises(ValueError):
        fe[80]([np.zeros(100, dtype=np.float32)], sampling_rate=8000)


@pytest.mark.parametrize("n_mels,B", [(80, 3), (128, 2), (80, 41)])
def test_tensor_core_dft_path_agrees_with_the_fft_kernel_and_the_oracle(fe, n_mels, B):
    """dwb_logmel_tc (the product path: windowed DFT as a tcgen05 GEMM with an fp16 hi/lo split, frames delivered by TMA through an
    overlapping-window tensor map) against dwb_logmel (shared-memory FFT) and the float64 oracle; B = 41 crosses a chunk boundary
    (37 utterances per chunk, double-buffered scratch, two internal streams)."""
    wav = lo.synthetic_waveforms(B, seed=11, ragged=True)
    wav[0] *= 0.003                                          # a very quiet clip: the hi/lo split must not lose it
    if B > 2:
        t = np.arange(lo.N_SAMPLES) / 16000.0
        wav[2] = (0.5 * np.sin(2 * np.pi * 440.0 * t) + 1e-4 * np.random.RandomState(0).randn(lo.N_SAMPLES)).astype(np.float32)
    w = torch.from_numpy(wav).cuda()
    tc = fe[n_mels].extract_device(w, impl="tc").cpu().numpy()
    fft = fe[n_mels].extract_device(w, impl="fft").cpu().numpy()
    assert tc.shape == (B, n_mels, 3000)
    assert np.abs(tc - fft).max() < 1e-4, np.abs(tc - fft).max()
    n_ref = min(B, 4)
    ref = lo.log_mel(wav[:n_ref], n_mels)
    err = np.abs(tc[:n_ref] - ref)
    assert err.max() < 1e-4 and np.quantile(err, 0.999) < ATOL, (err.max(), np.quantile(err, 0.999))
    # a second call on the same plan (stream fork / join state is reusable) is bitwise identical
    again = fe[n_mels].extract_device(w, impl="tc").cpu().numpy()
    assert np.array_equal(tc, again)
