"""One eager single-token decode step of the large-v3-shaped teacher (batch 32) between cudaProfilerStart/Stop."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
from distil_whisper_b200.modeling import DistilWhisperB200ForConditionalGeneration
from distil_whisper_b200 import generation, engine
torch.manual_seed(0)
with torch.device("cuda"):
    m = DistilWhisperB200ForConditionalGeneration(bench.TEACHER)
m = m.to(torch.bfloat16).eval()
B = 32
feats = bench.synthetic_batch(B, 8, 3, bench.TEACHER, device="cuda")["input_features"]
with torch.no_grad():
    enc, S, _ = engine.run_encoder(m, feats, None)
    ses = generation.DecodeSession(m, B, 128, S, use_graph=False)
    prompt = torch.tensor([[50258, 50259, 50360, 50364]], device="cuda").expand(B, -1).contiguous()
    ses.prepare(enc, prompt, None, None, 10 ** 6, 50256)
    for _ in range(40):
        ses._step(*ses.params)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ses._step(*ses.params)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
