"""CPU oracle for the Whisper log-mel feature extractor.  TEST INFRASTRUCTURE ONLY (see whisper_oracle.py).

numpy restatement of
  * HF:models/whisper/feature_extraction_whisper.py:135-164  _torch_extract_fbank_features
    (hann(400, periodic) ; STFT n_fft 400, hop 160, center/reflect -> 3001 frames, last dropped ;
     |.|^2 ; mel_filters.T @ ; log10(clamp 1e-10) ; max(x, per-utterance max - 8) ; (x + 4) / 4)
  * HF:audio_utils.py:263-296,299-340  slaney hertz<->mel ; :356-375 triangular bank ; :453-544 mel_filter_bank
    (norm="slaney", mel_scale="slaney", 0..8000 Hz, 201 bins -> n_mels)
  * HF:models/whisper/feature_extraction_whisper.py:281-296  pad / truncate to 480000 samples with zeros

Pinned by tests/golden/logmel.npz (made by oracle/gen_golden.py from the installed HF extractor).
Arithmetic is float64 inside the DFT and float32 at the HF rounding points that matter (power, mel, log).
"""
from __future__ import annotations

import numpy as np

N_FFT = 400
HOP = 160
N_SAMPLES = 480000
N_FRAMES = 3000
N_FREQ = N_FFT // 2 + 1
SAMPLING_RATE = 16000


def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    mels = 3.0 * f / 200.0
    logstep = 27.0 / np.log(6.4)
    log_region = f >= 1000.0
    out = mels.copy()
    out[log_region] = 15.0 + np.log(f[log_region] / 1000.0) * logstep
    return out


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f = 200.0 * m / 3.0
    logstep = np.log(6.4) / 27.0
    log_region = m >= 15.0
    out = f.copy()
    out[log_region] = 1000.0 * np.exp(logstep * (m[log_region] - 15.0))
    return out


def mel_filter_bank(n_mels: int = 80, n_freq: int = N_FREQ, sr: int = SAMPLING_RATE,
                    fmin: float = 0.0, fmax: float = 8000.0) -> np.ndarray:
    """[n_freq, n_mels] float64 slaney-scale, slaney-normalised triangular bank (HF:audio_utils.py:453-544)."""
    mel_pts = np.linspace(_hz_to_mel_slaney(np.array([fmin]))[0], _hz_to_mel_slaney(np.array([fmax]))[0], n_mels + 2)
    filter_freqs = _mel_to_hz_slaney(mel_pts)
    fft_freqs = np.linspace(0, sr // 2, n_freq)
    diff = np.diff(filter_freqs)
    slopes = filter_freqs[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / diff[:-1]
    up = slopes[:, 2:] / diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    enorm = 2.0 / (filter_freqs[2: n_mels + 2] - filter_freqs[:n_mels])
    return fb * enorm[None, :]


def hann_periodic(n: int = N_FFT) -> np.ndarray:
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)).astype(np.float32)


def pad_or_trim(wavs, n: int = N_SAMPLES) -> np.ndarray:
    out = np.zeros((len(wavs), n), dtype=np.float32)
    for i, w in enumerate(wavs):
        w = np.asarray(w, dtype=np.float32).reshape(-1)[:n]
        out[i, : len(w)] = w
    return out


def log_mel(waveforms: np.ndarray, n_mels: int = 80, chunk: int = 8) -> np.ndarray:
    """[B, 480000] float32 -> [B, n_mels, 3000] float32."""
    wav = np.asarray(waveforms, dtype=np.float32)
    assert wav.ndim == 2 and wav.shape[1] == N_SAMPLES
    B = wav.shape[0]
    win = hann_periodic().astype(np.float64)
    fb = mel_filter_bank(n_mels).astype(np.float32)                  # HF casts the bank to float32
    out = np.empty((B, n_mels, N_FRAMES), dtype=np.float32)
    idx = (np.arange(N_FRAMES)[:, None] * HOP + np.arange(N_FFT)[None, :])
    for b0 in range(0, B, chunk):
        w = wav[b0: b0 + chunk]
        padded = np.pad(w, ((0, 0), (N_FFT // 2, N_FFT // 2)), mode="reflect").astype(np.float64)
        frames = padded[:, idx] * win                                  # [b, 3000, 400]  (frame 3000 dropped)
        spec = np.fft.rfft(frames, axis=-1)                            # [b, 3000, 201]
        power = (spec.real ** 2 + spec.imag ** 2).astype(np.float32)
        mel = np.einsum("fm,btf->bmt", fb, power, optimize=True).astype(np.float32)
        logspec = np.log10(np.maximum(mel, 1e-10)).astype(np.float32)
        mx = logspec.reshape(logspec.shape[0], -1).max(axis=1)[:, None, None]
        logspec = np.maximum(logspec, mx - 8.0)
        out[b0: b0 + chunk] = (logspec + 4.0) / 4.0
    return out


def synthetic_waveforms(batch: int, seed: int, ragged: bool = True) -> np.ndarray:
    """randn*0.1 noise; with ragged=True ~30% of clips are hard-zero after a
    random length (exercises the 1e-10 clamp and the max-8 floor)."""
    rs = np.random.RandomState(seed)
    w = (0.1 * rs.randn(batch, N_SAMPLES)).astype(np.float32)
    if ragged:
        for b in range(batch):
            if rs.rand() < 0.3:
                n = int(rs.randint(16000, N_SAMPLES - 16000))
                w[b, n:] = 0.0
    return w
