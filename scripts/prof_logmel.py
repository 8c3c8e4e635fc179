import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from distil_whisper_b200.feature_extraction import WhisperFeatureExtractorB200
fe = WhisperFeatureExtractorB200(80)
wav = torch.randn((37, 480000), device="cuda") * 0.1
for _ in range(3):
    fe.extract_device(wav)
torch.cuda.synchronize()
