"""Run one kernel a few times (for ncu).  usage: prof_one.py gemm M N K [act] | attn B H Sq Sk causal tc | logmel B"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from distil_whisper_b200 import ops  # noqa: E402

kind = sys.argv[1]
if kind == "gemm":
    M, N, K = map(int, sys.argv[2:5])
    act = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    a = torch.randn((M, K), device="cuda").bfloat16()
    b = torch.randn((N, K), device="cuda").bfloat16()
    bias = torch.zeros(N, device="cuda")
    c = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(a, b, bias=bias, act=act, out=c)
elif kind == "attn":
    B, H, Sq, Sk, causal, tc = map(int, sys.argv[2:8])
    d = H * 64
    qkv = torch.randn((B * Sq, 3 * d), device="cuda").bfloat16()
    o = torch.empty((B * Sq, d), device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.attention_fwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], B, H, Sq, Sk, bool(causal), out=o, use_tc=bool(tc))
elif kind == "attn_bwd":
    B, H, Sq, Sk, causal, tc = map(int, sys.argv[2:8])
    d = H * 64
    q = torch.randn((B * Sq, d), device="cuda").bfloat16()
    k = torch.randn((B * Sk, d), device="cuda").bfloat16()
    v = torch.randn((B * Sk, d), device="cuda").bfloat16()
    do = torch.randn((B * Sq, d), device="cuda").bfloat16()
    o, lse = ops.attention_fwd(q, k, v, B, H, Sq, Sk, bool(causal), use_tc=True)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    for _ in range(3):
        ops.attention_bwd(q, k, v, o, do, lse, B, H, Sq, Sk, bool(causal), dq, dk, dv, use_tc=bool(tc))
elif kind == "logmel":
    from distil_whisper_b200.feature_extraction import WhisperFeatureExtractorB200
    B = int(sys.argv[2])
    fe = WhisperFeatureExtractorB200()
    wav = torch.randn((B, 480000), device="cuda") * 0.1
    for _ in range(3):
        fe.extract_device(wav)
torch.cuda.synchronize()
