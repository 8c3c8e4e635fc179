"""Log-mel feature extractor with the call surface of transformers.WhisperFeatureExtractor, computed by the
sm_100a kernel behind dwb_logmel (include/dwb.h).

Mirrors HF:models/whisper/feature_extraction_whisper.py:189-342 (`__call__`: list of 1-D float arrays, pad /
truncate to 30 s with zeros, -> {"input_features": [B, n_mels, 3000] float32}) as used by
ref:training/run_distillation.py:1176-1177 and :1234-1235.  The mel bank follows HF:audio_utils.py:453-544 with the
extractor's arguments (norm="slaney", mel_scale="slaney", 0..8000 Hz).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _abi


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    out = 3.0 * f / 200.0
    log_region = f >= 1000.0
    out[log_region] = 15.0 + np.log(f[log_region] / 1000.0) * (27.0 / np.log(6.4))
    return out


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    out = 200.0 * m / 3.0
    log_region = m >= 15.0
    out[log_region] = 1000.0 * np.exp((np.log(6.4) / 27.0) * (m[log_region] - 15.0))
    return out


def slaney_mel_filter_bank(num_frequency_bins: int, num_mel_filters: int, min_frequency: float, max_frequency: float,
                           sampling_rate: int) -> np.ndarray:
    """[num_frequency_bins, num_mel_filters] triangular bank, slaney scale + slaney area normalisation."""
    mel_pts = np.linspace(_hz_to_mel(np.array([min_frequency]))[0], _hz_to_mel(np.array([max_frequency]))[0],
                          num_mel_filters + 2)
    filter_freqs = _mel_to_hz(mel_pts)
    fft_freqs = np.linspace(0, sampling_rate // 2, num_frequency_bins)
    diff = np.diff(filter_freqs)
    slopes = filter_freqs[None, :] - fft_freqs[:, None]
    bank = np.maximum(0.0, np.minimum(-slopes[:, :-2] / diff[:-1], slopes[:, 2:] / diff[1:]))
    bank *= (2.0 / (filter_freqs[2: num_mel_filters + 2] - filter_freqs[:num_mel_filters]))[None, :]
    return bank


class WhisperFeatureExtractorB200:
    model_input_names = ["input_features"]

    def __init__(self, feature_size=80, sampling_rate=16000, hop_length=160, chunk_length=30, n_fft=400, padding_value=0.0,
                 **kwargs):
        if n_fft != 400 or hop_length != 160:
            raise ValueError("the B200 log-mel kernel is specialised for Whisper's n_fft=400 / hop_length=160")
        self.feature_size = feature_size
        self.sampling_rate = sampling_rate
        self.hop_length = hop_length
        self.chunk_length = chunk_length
        self.n_fft = n_fft
        self.padding_value = padding_value
        self.n_samples = chunk_length * sampling_rate
        self.nb_max_frames = self.n_samples // hop_length
        self.mel_filters = slaney_mel_filter_bank(1 + n_fft // 2, feature_size, 0.0, 8000.0, sampling_rate)
        self._plan = None
        self._plan_tc = None
        self._ws = None

    # -- (de)serialisation: ref:training/run_distillation.py:1071, :1641, :1754, :1783 call feature_extractor.save_pretrained ----
    def to_dict(self):
        return {"feature_extractor_type": "WhisperFeatureExtractor", "processor_class": "WhisperProcessor",
                "feature_size": self.feature_size, "sampling_rate": self.sampling_rate, "hop_length": self.hop_length,
                "chunk_length": self.chunk_length, "n_fft": self.n_fft, "padding_value": self.padding_value, "padding_side": "right",
                "n_samples": self.n_samples, "nb_max_frames": self.nb_max_frames, "return_attention_mask": False}

    def save_pretrained(self, save_directory, **kwargs):
        """Writes preprocessor_config.json in transformers' layout (loadable by WhisperFeatureExtractor.from_pretrained)."""
        import json
        import os
        os.makedirs(save_directory, exist_ok=True)
        path = os.path.join(save_directory, "preprocessor_config.json")
        with open(path, "w") as f:
            json.dump(self.to_dict(), f, indent=2, sort_keys=True)
        return [path]

    @classmethod
    def from_pretrained(cls, directory, **kwargs):
        import json
        import os
        with open(os.path.join(directory, "preprocessor_config.json")) as f:
            cfg = json.load(f)
        keys = ("feature_size", "sampling_rate", "hop_length", "chunk_length", "n_fft", "padding_value")
        return cls(**{k: cfg[k] for k in keys if k in cfg})

    # -- device plan ------------------------------------------------------------------------------------------
    def _get_plan(self):
        if self._plan is None:
            filt = np.ascontiguousarray(self.mel_filters.astype(np.float32))
            plan = C.c_void_p()
            _abi.call("dwb_logmel_plan_create", filt.ctypes.data_as(C.c_void_p), filt.shape[0], filt.shape[1], C.byref(plan))
            self._plan = plan
        return self._plan

    def __del__(self):
        try:
            if getattr(self, "_plan", None) is not None:
                _abi.call("dwb_logmel_plan_destroy", self._plan)
            if getattr(self, "_plan_tc", None) is not None:
                _abi.call("dwb_logmel_tc_plan_destroy", self._plan_tc)
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass

    def _get_plan_tc(self):
        if self._plan_tc is None:
            filt = np.ascontiguousarray(self.mel_filters.astype(np.float32))
            plan = C.c_void_p()
            _abi.call("dwb_logmel_tc_plan_create", filt.ctypes.data_as(C.c_void_p), filt.shape[0], filt.shape[1], C.byref(plan))
            self._plan_tc = plan
        return self._plan_tc

    def extract_device(self, wav: torch.Tensor, out: torch.Tensor | None = None, impl: str = "tc") -> torch.Tensor:
        """wav: CUDA fp32 [B, n_samples] (already padded / truncated).  Returns CUDA fp32 [B, n_mels, n_samples/160].
        impl "tc": the DFT as a tcgen05 GEMM (dwb_logmel_tc, the product path); "fft": the shared-memory FFT kernel
        (dwb_logmel), kept as an on-device cross-check."""
        if not (wav.is_cuda and wav.dtype == torch.float32 and wav.dim() == 2 and wav.is_contiguous()):
            raise ValueError("extract_device expects a contiguous CUDA float32 [B, n_samples] tensor")
        B, n = wav.shape
        if out is None:
            out = torch.empty((B, self.feature_size, n // self.hop_length), dtype=torch.float32, device=wav.device)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        if impl == "fft":
            _abi.call("dwb_logmel", self._get_plan(), C.c_void_p(wav.data_ptr()), B, n, C.c_void_p(out.data_ptr()), stream)
            return out
        if impl != "tc":
            raise ValueError(f"unknown log-mel implementation {impl!r}")
        need = int(_abi.call("dwb_logmel_tc_workspace_bytes", B, n))
        if self._ws is None or self._ws.numel() < need or self._ws.device != wav.device:
            self._ws = torch.empty((need,), dtype=torch.uint8, device=wav.device)
        _abi.call("dwb_logmel_tc", self._get_plan_tc(), C.c_void_p(wav.data_ptr()), B, n, C.c_void_p(out.data_ptr()),
                  C.c_void_p(self._ws.data_ptr()), stream)
        return out

    # -- HF-compatible surface -------------------------------------------------------------------------------
    def pad_or_trim(self, raw_speech) -> np.ndarray:
        if isinstance(raw_speech, np.ndarray) and raw_speech.ndim == 2:
            raw_speech = list(raw_speech)
        elif isinstance(raw_speech, np.ndarray) or (len(raw_speech) and np.isscalar(raw_speech[0])):
            raw_speech = [np.asarray(raw_speech)]
        buf = np.full((len(raw_speech), self.n_samples), self.padding_value, dtype=np.float32)
        for i, w in enumerate(raw_speech):
            w = np.asarray(w, dtype=np.float32).reshape(-1)[: self.n_samples]
            buf[i, : len(w)] = w
        return buf

    def __call__(self, raw_speech, sampling_rate=None, return_tensors=None, device="cuda", **kwargs):
        if sampling_rate is not None and sampling_rate != self.sampling_rate:
            raise ValueError(
                f"The model corresponding to this feature extractor was trained using a sampling rate of "
                f"{self.sampling_rate}. Please make sure that the provided `raw_speech` input was sampled with "
                f"{self.sampling_rate} and not {sampling_rate}.")
        host = torch.from_numpy(self.pad_or_trim(raw_speech))
        feats = self.extract_device(host.pin_memory().to(device, non_blocking=True))
        if return_tensors == "pt":
            return {"input_features": feats}
        arr = feats.cpu().numpy()
        if return_tensors == "np":
            return {"input_features": arr}
        return {"input_features": [a for a in arr]}

    def pad(self, processed_features, padding=True, return_tensors=None, **kwargs):
        """Collator hook (ref:training/run_distillation.py:447-451): features are already fixed-length, so stack."""
        feats = processed_features["input_features"]
        if isinstance(feats, torch.Tensor):
            stacked = feats
        else:
            stacked = torch.stack([torch.as_tensor(np.asarray(f), dtype=torch.float32) for f in feats])
        if return_tensors == "np":
            stacked = stacked.cpu().numpy()
        return {"input_features": stacked}
