"""Do the optimiser-tail kernels (side stream) actually run concurrently with the persistent tcgen05 GEMM / attention kernels?"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from distil_whisper_b200 import _abi, ops  # noqa: E402

dev = "cuda"
M, N, K = 48000, 3840, 1280
a = torch.randn((M, K), device=dev).bfloat16()
b = torch.randn((N, K), device=dev).bfloat16()
c = torch.empty((M, N), device=dev, dtype=torch.bfloat16)
n = 120_000_000
p, g, m, v = [torch.zeros(n, device=dev) for _ in range(4)]
ss = torch.zeros(1, device=dev)
side = torch.cuda.Stream()
B, H, S = 32, 20, 1500
qkv = torch.randn((B * S, 3 * H * 64), device=dev).bfloat16()
o = torch.empty((B * S, H * 64), device=dev, dtype=torch.bfloat16)


def gemms(k=40):
    for _ in range(k):
        ops.gemm(a, b, out=c)


def attns(k=20):
    d = H * 64
    for _ in range(k):
        ops.attention_fwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], B, H, S, S, False, out=o, need_lse=False, use_tc=True)


def tail():
    ops.grad_sumsq(g, ss)
    ops.adamw_step(p, g, m, v, None, 1e-4, 0.9, 0.999, 1e-8, 0.0, 1, ss, 1.0)


def timed(fn):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e)


def both(main_fn):
    def f():
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(side):
            side.wait_event(ev)
            tail()
            done = torch.cuda.Event()
            done.record(side)
        main_fn()
        torch.cuda.current_stream().wait_event(done)
    return f


for grid in (0, 48, 16):
    _abi.call("dwb_set_tail_grid", grid)
    for name, fn in (("gemm x40", gemms), ("attn x20", attns)):
        fn(); tail()
        t_main, t_tail, t_both = timed(fn), timed(tail), timed(both(fn))
        print(f"tail_grid {grid:3d} {name}: main {t_main:.2f} ms, tail alone {t_tail:.2f} ms, concurrent {t_both:.2f} ms "
              f"(sum {t_main + t_tail:.2f}, max {max(t_main, t_tail):.2f})")
