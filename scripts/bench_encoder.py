"""BASELINE.json configs[4]: distil-medium.en encoder-only forward + backward, batch 64, S = 1500 (218.5 TFLOP / step)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from distil_whisper_b200 import engine, ops  # noqa: E402
from distil_whisper_b200.modeling import DistilWhisperB200ForConditionalGeneration  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = dict(vocab_size=51864, num_mel_bins=80, d_model=1024, encoder_layers=24, encoder_attention_heads=16, encoder_ffn_dim=4096,
           decoder_layers=2, decoder_attention_heads=16, decoder_ffn_dim=4096, max_source_positions=1500, max_target_positions=448,
           pad_token_id=50256, decoder_start_token_id=50257)
dev = torch.device("cuda", 0)
with torch.device(dev):
    model = DistilWhisperB200ForConditionalGeneration(cfg)
enc = model.model.encoder
st = engine.state_of(enc)
feats = (0.5 * torch.randn((B, 80, 3000), device=dev)).clamp_(-1, 1.5)
denc = torch.randn((B * 1500, 1024), device=dev).bfloat16() * 1e-3


def step():
    out, ctx = engine.encoder_forward(st, feats, save=True)
    engine.encoder_backward(st, ctx, denc)
    for p in enc.parameters():
        if p.grad is not None:
            p.grad = None


for _ in range(2):
    step()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 3
s.record()
for _ in range(n):
    step()
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / n
fwd_gf_per_utt = 1138.1
tf = 3 * fwd_gf_per_utt * B / 1e3
res = dict(config="distil-medium.en encoder fwd+bwd", B=B, ms_per_step=round(ms, 2), tflop_per_step=round(tf, 1),
           tflops=round(tf / ms * 1e3, 1), peak_mem_gb=round(torch.cuda.max_memory_allocated() / 2**30, 1))
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bench_encoder.json", "w"))
