import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from distil_whisper_b200.feature_extraction import WhisperFeatureExtractorB200
from oracle import logmel_oracle as lo
for n_mels in (80, 128):
    fe = WhisperFeatureExtractorB200(n_mels)
    wav = lo.synthetic_waveforms(5, seed=3, ragged=True)
    wav[2] *= 0.001                                   # a very quiet clip
    t = np.arange(480000) / 16000.0
    wav[3] = (0.5 * np.sin(2 * np.pi * 440 * t) + 1e-4 * np.random.RandomState(0).randn(480000)).astype(np.float32)   # tone + faint noise floor
    w = torch.from_numpy(wav).cuda()
    a = fe.extract_device(w, impl="fft").cpu().numpy()
    b = fe.extract_device(w, impl="tc").cpu().numpy()
    ref = lo.log_mel(wav, n_mels)
    print(n_mels, "fft vs oracle", np.abs(a - ref).max(axis=(1, 2)), "tc vs oracle", np.abs(b - ref).max(axis=(1, 2)), "tc vs fft", np.abs(a - b).max())
fe = WhisperFeatureExtractorB200(80)
wav = torch.randn((1024, 480000), device="cuda") * 0.1
out = torch.empty((1024, 80, 3000), device="cuda")
for impl in ("fft", "tc"):
    for _ in range(2):
        fe.extract_device(wav, out=out, impl=impl)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        fe.extract_device(wav, out=out, impl=impl)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    print(impl, "ms per 1024 clips", round(ms, 3), "GB/s", round(1024 * 2.88e6 / ms / 1e6, 1))
