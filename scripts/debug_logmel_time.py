import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from distil_whisper_b200.feature_extraction import WhisperFeatureExtractorB200
fe = WhisperFeatureExtractorB200(80)
wav = torch.randn((1024, 480000), device="cuda") * 0.1
out = torch.empty((1024, 80, 3000), device="cuda")
for _ in range(2):
    fe.extract_device(wav, out=out)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    fe.extract_device(wav, out=out)
e.record(); torch.cuda.synchronize()
print("DWB_LOGMEL_DBG", os.environ.get("DWB_LOGMEL_DBG", "0"), "ms per 1024 clips", round(s.elapsed_time(e) / 5, 3))
