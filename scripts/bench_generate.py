"""Pseudo-labelling decode throughput (ref:training/run_pseudo_labelling.py:861-927): the large-v3-shaped teacher
transcribes a batch greedily with the KV-cached single-token CUDA graph.  Also times the distil-large-v3 student (eval loop)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench  # noqa: E402
from distil_whisper_b200.modeling import DistilWhisperB200ForConditionalGeneration  # noqa: E402

res = []
CASES = (("large-v3 teacher (bf16)", bench.TEACHER, 32), ("large-v3 teacher (bf16)", bench.TEACHER, 64),
         ("distil-large-v3 student (fp32 masters)", bench.STUDENT, 32))
if len(sys.argv) > 1:
    CASES = tuple(c for c in CASES if str(c[2]) in sys.argv[1:] and "teacher" in c[0])
for name, dims, B in CASES:
    torch.manual_seed(0)
    with torch.device("cuda"):
        m = DistilWhisperB200ForConditionalGeneration(dims)
    if "teacher" in name:
        m = m.to(torch.bfloat16)
    feats = bench.synthetic_batch(B, 8, 3, dims, device="cuda")["input_features"]
    n_new = 124
    kw = dict(max_new_tokens=n_new, eos_token_id=10 ** 6, decoder_input_ids=torch.tensor([[50258, 50259, 50360, 50364]], device="cuda").expand(B, -1))
    m.generate(feats, **kw)                      # builds the session + graph
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = m.generate(feats, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # decode-only time: encoder excluded
    from distil_whisper_b200 import engine
    t1 = time.perf_counter()
    engine.run_encoder(m, feats, None)
    torch.cuda.synchronize()
    t_enc = time.perf_counter() - t1
    steps = out.shape[1] - 1
    res.append(dict(model=name, batch=B, tokens_per_row=int(out.shape[1]), s_total=round(dt, 4), s_encoder=round(t_enc, 4),
                    ms_per_token_step=round((dt - t_enc) / steps * 1e3, 3), tokens_per_s=round(B * steps / dt, 1),
                    utterances_per_s=round(B / dt, 2)))
    print(res[-1], flush=True)
    del m
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bench_generate.json", "w"), indent=1)
