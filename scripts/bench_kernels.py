"""Micro-benchmarks of individual kernels (CUDA events, L2-flushing between iterations is implicit: operands > L2).
Usage: python scripts/bench_kernels.py [gemm] [attn]"""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from distil_whisper_b200 import ops  # noqa: E402


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def bench_gemm():
    out = []
    for (M, N, K, amn, bmn, act, f32) in [
        (48000, 1280, 1280, 0, 0, 0, 0), (48000, 3840, 1280, 0, 0, 0, 0), (48000, 5120, 1280, 0, 0, 1, 0),
        (48000, 1280, 5120, 0, 0, 0, 0), (48000, 2560, 1280, 0, 0, 0, 0), (4096, 1280, 1280, 0, 0, 0, 0),
        (4096, 5120, 1280, 0, 0, 1, 0), (4096, 51866, 1280, 0, 0, 0, 1), (4096, 1280, 51866, 0, 1, 0, 0),
        (1280, 1280, 4096, 1, 1, 0, 1), (51866, 1280, 4096, 1, 1, 0, 1), (8192, 8192, 8192, 0, 0, 0, 0),
        (4096, 3840, 1280, 0, 0, 0, 0), (4096, 1280, 5120, 0, 0, 0, 0), (5120, 1280, 4096, 1, 1, 0, 1),
    ]:
        def mk(r, c):
            return torch.randn((r, ops.round_up(c, 8)), device="cuda").bfloat16()[:, :c]
        a = mk(K, M) if amn else mk(M, K)
        b = mk(K, N) if bmn else mk(N, K)
        bias = torch.zeros(N, device="cuda") if not f32 else None
        ld = ops.round_up(N, 8)
        c = torch.empty((M, ld), device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)[:, :N]
        ms = timeit(lambda: ops.gemm(a, b, a_mn=bool(amn), b_mn=bool(bmn), bias=bias, act=act, out=c))
        ms1 = timeit(lambda: ops.gemm(a, b, a_mn=bool(amn), b_mn=bool(bmn), bias=bias, act=act, out=c, impl=3))
        ms2 = timeit(lambda: ops.gemm(a, b, a_mn=bool(amn), b_mn=bool(bmn), bias=bias, act=act, out=c, impl=2)) if N >= 256 else float("nan")
        aT = a.t() if amn else a
        bT = b if bmn else b.t()
        ms_t = timeit(lambda: torch.matmul(aT, bT))
        tf = 2.0 * M * N * K / ms / 1e9
        out.append(dict(M=M, N=N, K=K, a_mn=amn, b_mn=bmn, act=act, f32=f32, ms=round(ms, 4), tflops=round(tf, 1),
                        single_cta_tflops=round(2.0 * M * N * K / ms1 / 1e9, 1), cta_pair_tflops=round(2.0 * M * N * K / ms2 / 1e9, 1),
                        cublas_ms=round(ms_t, 4), cublas_tflops=round(2.0 * M * N * K / ms_t / 1e9, 1)))
        print(out[-1], flush=True)
    return out


def bench_attn():
    out = []
    for (B, H, Sq, Sk, causal, tc) in [(32, 20, 1500, 1500, 0, 0), (32, 20, 1500, 1500, 0, 1), (32, 20, 128, 1500, 0, 0), (32, 20, 128, 128, 1, 0)]:
        d = H * 64
        qkv = torch.randn((B * Sq, 3 * d), device="cuda").bfloat16()
        kv = torch.randn((B * Sk, 2 * d), device="cuda").bfloat16()
        q = qkv[:, :d]
        k, v = (qkv[:, d:2 * d], qkv[:, 2 * d:]) if Sq == Sk else (kv[:, :d], kv[:, d:])
        o = torch.empty((B * Sq, d), device="cuda", dtype=torch.bfloat16)
        try:
            ms = timeit(lambda: ops.attention_fwd(q, k, v, B, H, Sq, Sk, bool(causal), out=o, use_tc=bool(tc)))
        except Exception as ex:  # noqa: BLE001
            print("skip", (B, H, Sq, Sk, causal, tc), ex)
            continue
        fl = 4.0 * B * H * Sq * Sk * 64 * (0.5 if causal else 1.0)
        qq = q.reshape(B, Sq, H, 64).transpose(1, 2)
        kk = k.reshape(B, Sk, H, 64).transpose(1, 2)
        vv = v.reshape(B, Sk, H, 64).transpose(1, 2)
        ms_t = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qq, kk, vv, is_causal=bool(causal)))
        out.append(dict(B=B, H=H, Sq=Sq, Sk=Sk, causal=causal, tc=tc, ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1),
                        sdpa_ms=round(ms_t, 4), sdpa_tflops=round(fl / ms_t / 1e9, 1)))
        print(out[-1], flush=True)
    return out


def bench_attn_bwd():
    out = []
    for (B, H, Sq, Sk, causal) in [(32, 20, 1500, 1500, 0), (32, 20, 128, 1500, 0), (32, 20, 128, 128, 1)]:
        d = H * 64
        q = torch.randn((B * Sq, d), device="cuda").bfloat16()
        k = torch.randn((B * Sk, d), device="cuda").bfloat16()
        v = torch.randn((B * Sk, d), device="cuda").bfloat16()
        do = torch.randn((B * Sq, d), device="cuda").bfloat16()
        o, lse = ops.attention_fwd(q, k, v, B, H, Sq, Sk, bool(causal), use_tc=not causal)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        fl = 10.0 * B * H * Sq * Sk * 64 * (0.5 if causal else 1.0)
        row = dict(B=B, H=H, Sq=Sq, Sk=Sk, causal=causal)
        for name, tc in (("tc", True), ("mma_sync", False)):
            ms = timeit(lambda: ops.attention_bwd(q, k, v, o, do, lse, B, H, Sq, Sk, bool(causal), dq, dk, dv, use_tc=tc), iters=5)
            row[name + "_ms"] = round(ms, 4)
            row[name + "_tflops"] = round(fl / ms / 1e9, 1)
        out.append(row)
        print(row, flush=True)
    return out


def bench_logmel():
    """BASELINE.json configs[3]: 1024 x 480000-sample fp32 waveforms -> [1024, 80, 3000]; algorithmic bytes 2.88 MB / utterance."""
    from distil_whisper_b200.feature_extraction import WhisperFeatureExtractorB200
    out = []
    for n_mels, B in ((80, 1024), (128, 1024)):
        fe = WhisperFeatureExtractorB200(n_mels)
        wav = torch.randn((B, 480000), device="cuda") * 0.1
        dst = torch.empty((B, n_mels, 3000), device="cuda")
        ms = timeit(lambda: fe.extract_device(wav, out=dst), iters=5, warmup=3)
        byts = B * (480000 * 4 + n_mels * 3000 * 4)
        # the library path the reference reaches: torch.stft + matmul + log10 (HF _torch_extract_fbank_features on the GPU)
        win = torch.hann_window(400, device="cuda")
        filt = torch.from_numpy(fe.mel_filters.astype("float32")).cuda()

        def hf_like():
            st = torch.stft(wav[:256], 400, 160, window=win, return_complex=True)
            mag = st[..., :-1].abs() ** 2
            ls = torch.clamp(filt.T @ mag, min=1e-10).log10()
            mx = ls.amax(dim=(1, 2), keepdim=True)
            return (torch.maximum(ls, mx - 8.0) + 4.0) / 4.0
        ms_t = timeit(hf_like, iters=3, warmup=2) * (B / 256)
        out.append(dict(n_mels=n_mels, B=B, ms=round(ms, 4), gbps=round(byts / ms / 1e6, 1), utt_per_s=round(B / ms * 1e3),
                        torch_ms=round(ms_t, 3), torch_gbps=round(byts / ms_t / 1e6, 1)))
        print(out[-1], flush=True)
    return out


def bench_ln():
    out = []
    for rows, d in [(48000, 1280), (4096, 1280)]:
        x = torch.randn((rows, d), device="cuda")
        g, b = torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")
        bufs = [torch.randn((rows, d), device="cuda") for _ in range(3)]      # rotate inputs: 3 x 245 MB > L2
        it = [0]

        def run():
            it[0] += 1
            ops.add_layernorm(bufs[it[0] % 3], None, g, b, rows=rows, d=d, write_x=False, write_ln=True)
        ms = timeit(run, iters=12)
        nbytes = rows * d * 6
        out.append(dict(rows=rows, d=d, ms=round(ms, 4), gbps=round(nbytes / ms / 1e6, 1)))
        print(out[-1], flush=True)
    return out


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "attn", "logmel"]
    res = {}
    if "gemm" in which:
        res["gemm"] = bench_gemm()
    if "attn" in which:
        res["attn"] = bench_attn()
    if "attn_bwd" in which:
        res["attn_bwd"] = bench_attn_bwd()
    if "logmel" in which:
        res["logmel"] = bench_logmel()
    if "ln" in which:
        res["ln"] = bench_ln()
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bench_kernels.json", "w") as f:
        json.dump(res, f, indent=1)
