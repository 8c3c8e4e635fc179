"""Run-to-run determinism of individual kernels at the decoder's shapes (same inputs, two launches, bitwise / relative diff)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from distil_whisper_b200 import ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (a.norm() + 1e-30)), float((a != b).double().mean())


B, H, T, S = 32, 20, 128, 1500
d = H * 64
for name, (Sq, Sk, causal) in {"cross": (T, S, False), "self-causal": (T, T, True), "encoder": (S, S, False)}.items():
    Bx = B if name != "encoder" else 4
    q = torch.randn((Bx * Sq, d), device=dev).bfloat16()
    k = torch.randn((Bx * Sk, d), device=dev).bfloat16()
    v = torch.randn((Bx * Sk, d), device=dev).bfloat16()
    do = (torch.randn((Bx * Sq, d), device=dev) * 0.1).bfloat16()
    outs = []
    for rep in range(2):
        o, lse = ops.attention_fwd(q, k, v, Bx, H, Sq, Sk, causal, use_tc=True)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        ops.attention_bwd(q, k, v, o, do, lse, Bx, H, Sq, Sk, causal, dq, dk, dv, use_tc=True)
        torch.cuda.synchronize()
        outs.append((o.clone(), lse.clone(), dq.clone(), dk.clone(), dv.clone()))
    print(name, {n: rel(a, b) for n, a, b in zip(("o", "lse", "dq", "dk", "dv"), outs[0], outs[1])})
# LayerNorm backward
rows = B * T
x = torch.randn((rows, d), device=dev)
dy = (torch.randn((rows, d), device=dev) * 0.1).bfloat16()
g = torch.randn(d, device=dev)
_, ln, mu, rs = ops.add_layernorm(x, None, g, torch.zeros_like(g), rows=rows, d=d, save_stats=True, write_x=False)
outs = []
for rep in range(2):
    dg, db = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
    dx, dxb = ops.layernorm_bwd(dy, x, mu, rs, g, None, dg, db, rows=rows, d=d)
    torch.cuda.synchronize()
    outs.append((dx.clone(), dxb.clone(), dg.clone(), db.clone()))
print("ln_bwd", {n: rel(a, b) for n, a, b in zip(("dx", "dxb", "dg", "db"), outs[0], outs[1])})
# GEMMs: wgrad with accumulate (+ split-K), dgrad over the vocabulary
a = (torch.randn((48000, 1280), device=dev) * 0.1).bfloat16()
bm = (torch.randn((48000, 1280), device=dev) * 0.1).bfloat16()
outs = []
for rep in range(2):
    c = torch.zeros((1280, 1280), device=dev)
    ops.gemm(a, bm, a_mn=True, b_mn=True, out=c, accumulate=True)
    torch.cuda.synchronize()
    outs.append(c.clone())
print("wgrad split-K 1280x1280x48000", rel(outs[0], outs[1]))
dl = (torch.randn((4096, 51872), device=dev) * 0.01).bfloat16()[:, :51866]
E = (torch.randn((51866, 1280), device=dev) * 0.02).bfloat16()
outs = []
for rep in range(2):
    outs.append(ops.gemm(dl, E, b_mn=True).clone())
print("lm-head dgrad", rel(outs[0], outs[1]))
hf = (torch.randn((4096, 1280), device=dev)).bfloat16()
outs = []
for rep in range(2):
    c = torch.zeros((51866, 1280), device=dev)
    ops.gemm(dl, hf, a_mn=True, b_mn=True, out=c, accumulate=True)
    outs.append(c.clone())
print("lm-head wgrad", rel(outs[0], outs[1]))
outs = []
w = (torch.randn((5120, 1280), device=dev) * 0.02).bfloat16()
xx = torch.randn((4096, 1280), device=dev).bfloat16()
for rep in range(2):
    outs.append(ops.gemm(xx, w, bias=torch.zeros(5120, device=dev), act=1).clone())
print("fc1 gelu gemm", rel(outs[0], outs[1]))
