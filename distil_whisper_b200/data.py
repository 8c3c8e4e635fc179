"""On-GPU preprocessing feed for the KD step.

`DataCollatorSpeechSeq2SeqWithPaddingB200` keeps the call surface of the reference's collator
(ref:training/run_distillation.py:404-478: constructed with the processor, decoder_start_token_id, decoder_prev_token_id,
max_target_length; called with a list of feature dicts; returns {"input_features", "labels", "decoder_input_ids"}) and moves
its arithmetic to the device:
  * features given as RAW audio (`"input_values"` / `"audio"` 1-D float arrays, what `prepare_train_dataset` (ref :1167-1177)
    receives before it calls the feature extractor) are padded / truncated to 30 s into one pinned host buffer, copied once,
    and turned into log-mel features by dwb_logmel -- the mel tensor never exists on the host;
    features that are already log-mel (`"input_features"`) are stacked and copied (the reference's behaviour, ref :447-451);
  * label rows are padded on the host into one int64 matrix + lengths (that is list handling, not arithmetic), copied, and
    `decoder_input_ids` / `labels` (-100 on padding and on the prompt up to <|startoftranscript|>, ref :460-476) are built by
    dwb_collate_labels.
The result is a batch of CUDA tensors ready for DistillationStep.train_step / PipelinedTrainer.step.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Optional, Union

import numpy as np
import torch

from . import ops
from .feature_extraction import WhisperFeatureExtractorB200


def pad_label_rows(rows, pad_token_id: int, max_target_length: Optional[int] = None, padding: Union[bool, str] = "max_length"):
    """tokenizer.pad(label_features, max_length=..., padding=...) for right-padded Whisper tokenizers: -> (int64 [B, L1] pinned,
    int32 [B] lengths pinned).  padding 'max_length' pads to max_target_length, True / 'longest' to the longest row."""
    lens = [len(r) for r in rows]
    if padding == "max_length":
        if max_target_length is None:
            raise ValueError("padding='max_length' needs max_target_length")
        L1 = int(max_target_length)
        if max(lens) > L1:
            raise ValueError(f"a label row has {max(lens)} tokens > max_target_length {L1} (the reference filters these, ref :1247-1255)")
    elif padding in (True, "longest"):
        L1 = max(lens)
    else:
        raise ValueError(f"unsupported target padding {padding!r}")
    pin = torch.cuda.is_available()
    toks = torch.full((len(rows), L1), int(pad_token_id), dtype=torch.int64, pin_memory=pin)
    for i, r in enumerate(rows):
        toks[i, : lens[i]] = torch.as_tensor(np.asarray(r, dtype=np.int64))
    return toks, torch.tensor(lens, dtype=torch.int32, pin_memory=pin)


@dataclass
class DataCollatorSpeechSeq2SeqWithPaddingB200:
    processor: Any                       # WhisperProcessor-like (.feature_extractor, .tokenizer) or None
    decoder_start_token_id: int
    decoder_prev_token_id: int           # kept for signature parity (the reference's collator does not use it either)
    input_padding: Union[bool, str] = "max_length"
    target_padding: Union[bool, str] = "max_length"
    max_target_length: Optional[int] = None
    pad_token_id: Optional[int] = None   # defaults to processor.tokenizer.pad_token_id
    feature_extractor: Optional[WhisperFeatureExtractorB200] = None
    device: Union[str, torch.device] = "cuda"

    def __post_init__(self):
        if self.pad_token_id is None:
            tok = getattr(self.processor, "tokenizer", None)
            if tok is None or getattr(tok, "pad_token_id", None) is None:
                raise ValueError("pass pad_token_id (no processor.tokenizer.pad_token_id to read it from)")
            self.pad_token_id = int(tok.pad_token_id)
        if self.feature_extractor is None:
            fe = getattr(self.processor, "feature_extractor", None)
            n_mels = getattr(fe, "feature_size", 80)
            self.feature_extractor = fe if isinstance(fe, WhisperFeatureExtractorB200) else WhisperFeatureExtractorB200(n_mels)

    def _features(self, features):
        fe = self.feature_extractor
        first = features[0]
        raw_key = next((k for k in ("input_values", "audio") if k in first), None)
        if raw_key is not None:
            clips = [f[raw_key]["array"] if isinstance(f[raw_key], dict) else f[raw_key] for f in features]
            host = torch.from_numpy(fe.pad_or_trim(clips))
            if torch.cuda.is_available():
                host = host.pin_memory()
            return fe.extract_device(host.to(self.device, non_blocking=True))          # [B, n_mels, 3000] fp32, device only
        feats = [f["input_features"] for f in features]
        if isinstance(feats[0], torch.Tensor) and feats[0].is_cuda:
            return torch.stack(feats).to(torch.float32)
        host = torch.from_numpy(np.stack([np.asarray(x, dtype=np.float32) for x in feats]))
        if torch.cuda.is_available():
            host = host.pin_memory()
        return host.to(self.device, non_blocking=True)

    def __call__(self, features):
        batch = {"input_features": self._features(features)}
        toks, lens = pad_label_rows([f["labels"] for f in features], self.pad_token_id, self.max_target_length, self.target_padding)
        dec_in, labels = ops.collate_labels(toks.to(self.device, non_blocking=True), lens.to(self.device, non_blocking=True),
                                            self.decoder_start_token_id)
        batch["labels"] = labels
        batch["decoder_input_ids"] = dec_in
        return batch
