"""Tensor-facing wrappers over the C ABI (include/dwb.h).  torch is used for device memory and streams only:
every arithmetic op below runs in a kernel from libdwb.so, and every call raises if the library is absent."""
from __future__ import annotations

import ctypes as C

import torch

from . import _abi

BF16 = torch.bfloat16
F32 = torch.float32
HEAD_DIM = 64
# bench.py sets this to a list to time every tcgen05 GEMM launch with CUDA events: (start, end, 2*M*N*K) per launch
GEMM_PROFILE = None


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def _check2d(t, dtype, name):
    if t.dim() != 2 or t.stride(1) != 1 or t.dtype != dtype or not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA {dtype} matrix with unit column stride, got "
                         f"{tuple(t.shape)} {t.dtype} strides {t.stride()} on {t.device}")


def gemm(a, b, *, a_mn=False, b_mn=False, bias=None, act=0, out=None, out_dtype=BF16, alpha=1.0, accumulate=False,
         impl=0):
    """out[M,N] = act(alpha * A . B^T + bias).  a: [M,K] (or [K,M] when a_mn); b: [N,K] (or [K,N] when b_mn)."""
    _check2d(a, BF16, "gemm A")
    _check2d(b, BF16, "gemm B")
    K, M = (a.shape if a_mn else (a.shape[1], a.shape[0]))
    Kb, N = (b.shape if b_mn else (b.shape[1], b.shape[0]))
    if K != Kb:
        raise ValueError(f"gemm: reduction dims differ ({K} vs {Kb})")
    if out is None:
        ld = round_up(N, 8 if out_dtype == BF16 else 4)
        buf = torch.empty((M, ld), dtype=out_dtype, device=a.device)
        out = buf[:, :N]
    _check2d(out, out.dtype, "gemm C")
    if out.shape != (M, N) or out.dtype not in (BF16, F32):
        raise ValueError(f"gemm: bad output {tuple(out.shape)} {out.dtype} for M={M} N={N}")
    if bias is not None and (bias.dtype != F32 or bias.numel() != N or not bias.is_contiguous()):
        raise ValueError("gemm: bias must be contiguous fp32 [N]")
    prof = GEMM_PROFILE
    if prof is not None and impl == 0:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    _abi.call("dwb_gemm_bf16", _ptr(a), a.stride(0), int(a_mn), _ptr(b), b.stride(0), int(b_mn), _ptr(out), out.stride(0),
              int(out.dtype == F32), M, N, K, _ptr(bias), int(act), float(alpha), int(accumulate), int(impl), _stream())
    if prof is not None and impl == 0:
        ev1.record()
        prof.append((ev0, ev1, 2.0 * M * N * K))
    return out


def attention_fwd(q, k, v, B, H, Sq, Sk, causal, out=None, need_lse=True, use_tc=False):
    """q: [B*Sq, >=H*64] view, k/v: [B*Sk, .] views (unit column stride).  Returns (o [B*Sq, H*64] bf16, lse [B,H,Sq])."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _check2d(t, BF16, "attention " + n)
    if out is None:
        out = torch.empty((B * Sq, H * HEAD_DIM), dtype=BF16, device=q.device)
    lse = torch.empty((B, H, Sq), dtype=F32, device=q.device) if need_lse else None
    fn = "dwb_attention_fwd_tc" if use_tc else "dwb_attention_fwd"
    _abi.call(fn, _ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0), _ptr(out), out.stride(0), _ptr(lse),
              B, H, Sq, Sk, HEAD_DIM, int(causal), HEAD_DIM ** -0.5, _stream())
    return out, lse


USE_TC_ATTENTION_BWD = True     # tcgen05 attention backward; False -> mma.sync kernel (cross-check)


def attention_bwd(q, k, v, o, dout, lse, B, H, Sq, Sk, causal, dq_out, dk_out, dv_out, use_tc=None):
    """Writes dq_out/dk_out/dv_out (bf16 views with the layouts of q/k/v)."""
    dev = q.device
    delta = torch.empty((B * H * Sq,), dtype=F32, device=dev)
    dq_acc = torch.empty((B * Sq, H * HEAD_DIM), dtype=F32, device=dev)
    use_tc = USE_TC_ATTENTION_BWD if use_tc is None else use_tc
    _abi.call("dwb_attention_bwd_tc" if use_tc else "dwb_attention_bwd", _ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0), _ptr(o), o.stride(0),
              _ptr(dout), dout.stride(0), _ptr(lse), _ptr(delta), _ptr(dq_acc), _ptr(dk_out), dk_out.stride(0), _ptr(dv_out),
              dv_out.stride(0), B, H, Sq, Sk, HEAD_DIM, int(causal), HEAD_DIM ** -0.5, _stream())
    cast_f32_to_bf16(dq_acc, dq_out)
    return dq_out, dk_out, dv_out


def add_layernorm(x_in, y, gamma, beta, *, rows, d, x_rows_mod=0, write_x=True, write_ln=True, save_stats=False, eps=1e-5,
                  x_out=None):
    """x_new = x_in (+ y); returns (x_new fp32 or None, ln bf16 or None, mean, rstd)."""
    dev = x_in.device
    if write_x and x_out is None:
        x_out = torch.empty((rows, d), dtype=F32, device=dev)
    ln = torch.empty((rows, d), dtype=BF16, device=dev) if write_ln else None
    mean = torch.empty((rows,), dtype=F32, device=dev) if save_stats else None
    rstd = torch.empty((rows,), dtype=F32, device=dev) if save_stats else None
    _abi.call("dwb_add_layernorm", _ptr(x_in), int(x_rows_mod), _ptr(y), _ptr(gamma), _ptr(beta), _ptr(x_out if write_x else None),
              _ptr(ln), _ptr(mean), _ptr(rstd), rows, d, float(eps), _stream())
    return (x_out if write_x else None), ln, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, dres, dgamma, dbeta, *, rows, d, want_bf16=True):
    dev = x.device
    dx = torch.empty((rows, d), dtype=F32, device=dev)
    dxb = torch.empty((rows, d), dtype=BF16, device=dev) if want_bf16 else None
    _abi.call("dwb_layernorm_bwd", _ptr(dy), _ptr(x), _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(dres), _ptr(dx), _ptr(dxb),
              _ptr(dgamma), _ptr(dbeta), rows, d, _stream())
    return dx, dxb


def cast_f32_to_bf16(src, dst=None, scale=1.0):
    s2 = src if src.dim() == 2 else src.reshape(1, -1) if src.dim() == 1 else src.reshape(-1, src.shape[-1])
    if dst is None:
        dst = torch.empty(src.shape, dtype=BF16, device=src.device)
    d2 = dst if dst.dim() == 2 else dst.reshape(1, -1) if dst.dim() == 1 else dst.reshape(-1, dst.shape[-1])
    _check2d(s2, F32, "cast src")
    _check2d(d2, BF16, "cast dst")
    _abi.call("dwb_cast_f32_to_bf16", _ptr(s2), s2.stride(0), _ptr(d2), d2.stride(0), s2.shape[0], s2.shape[1], float(scale), _stream())
    return dst


def cast_bf16_to_f32(src, dst=None):
    s2 = src if src.dim() == 2 else src.reshape(-1, src.shape[-1])
    if dst is None:
        dst = torch.empty(src.shape, dtype=F32, device=src.device)
    d2 = dst if dst.dim() == 2 else dst.reshape(-1, dst.shape[-1])
    _abi.call("dwb_cast_bf16_to_f32", _ptr(s2), s2.stride(0), _ptr(d2), d2.stride(0), s2.shape[0], s2.shape[1], _stream())
    return dst


def conv_weight_to_kc(w):
    O, Cc, k = w.shape
    assert k == 3 and w.dtype == F32 and w.is_contiguous()
    out = torch.empty((O, 3 * Cc), dtype=BF16, device=w.device)
    _abi.call("dwb_conv_weight_to_kc_bf16", _ptr(w), _ptr(out), O, Cc, _stream())
    return out


def conv_wgrad_kc_to_ck(g, O, Cc):
    dw = torch.empty((O, Cc, 3), dtype=F32, device=g.device)
    _abi.call("dwb_conv_wgrad_kc_to_ck", _ptr(g), _ptr(dw), O, Cc, 0, _stream())
    return dw


def im2col_conv1(mel):
    B, Cc, L = mel.shape
    assert mel.dtype == F32 and mel.is_contiguous()
    ld = round_up(3 * Cc, 8)
    out = torch.empty((B * L, ld), dtype=BF16, device=mel.device)
    _abi.call("dwb_im2col_conv1", _ptr(mel), _ptr(out), B, Cc, L, ld, _stream())
    return out


def im2col_conv2(x, B, L, d):
    out = torch.empty((B * (L // 2), 3 * d), dtype=BF16, device=x.device)
    _abi.call("dwb_im2col_conv2", _ptr(x), _ptr(out), B, L, d, _stream())
    return out


def col2im_conv2_gelu_bwd(g, pre1, B, L, d):
    out = torch.empty((B * L, d), dtype=BF16, device=g.device)
    _abi.call("dwb_col2im_conv2_gelu_bwd", _ptr(g), _ptr(pre1), _ptr(out), B, L, d, _stream())
    return out


def embed_fwd(ids, E, P, B, T, d, vocab):
    assert ids.dtype == torch.int64 and ids.is_contiguous() and E.dtype == P.dtype
    x = torch.empty((B * T, d), dtype=F32, device=ids.device)
    _abi.call("dwb_embed_fwd", _ptr(ids), _ptr(E), _ptr(P), int(E.dtype == F32), _ptr(x), B, T, d, vocab, _stream())
    return x


def embed_bwd(ids, dx, dE, dP, B, T, d, vocab, padding_idx):
    _abi.call("dwb_embed_bwd", _ptr(ids), _ptr(dx), _ptr(dE), _ptr(dP), B, T, d, vocab, int(padding_idx), _stream())


import os as _os
_SKINNY = _os.environ.get("DWB_SKINNY_GEMM", "1") != "0"      # A/B switch for the decode-step projections


def gemm_small_m(a, b, *, bias=None, act=0, out_dtype=BF16):
    """out[M,N] = act(A . B^T + bias) for the decode step (M = batch rows).  Batch sizes of 16 / 32 / 48 / 64 take the weight-streaming
    skinny kernel (dwb_gemm_skinny_bf16); anything else goes through the tcgen05 GEMM."""
    M, K = a.shape
    N = b.shape[0]
    if not (_SKINNY and M in (16, 32, 48, 64) and N % 8 == 0 and K % 256 == 0 and a.stride(0) % 8 == 0 and b.stride(0) % 8 == 0
            and a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0):
        return gemm(a, b, bias=bias, act=act, out_dtype=out_dtype)
    _check2d(a, BF16, "gemm_small_m A")
    _check2d(b, BF16, "gemm_small_m B")
    out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    _abi.call("dwb_gemm_skinny_bf16", _ptr(a), a.stride(0), _ptr(b), b.stride(0), _ptr(out), out.stride(0), int(out_dtype == F32), M, N, K,
              _ptr(bias), int(act), _stream())
    return out


def embed_decode(seq, pos_dev, E, P, d, vocab):
    """x[b] = E[seq[b, pos]] + P[pos] (pos read on the device) -> fp32 [B, d]."""
    assert seq.dtype == torch.int64 and seq.dim() == 2 and seq.stride(1) == 1 and pos_dev.dtype == torch.int32 and E.dtype == P.dtype
    B = seq.shape[0]
    x = torch.empty((B, d), dtype=F32, device=seq.device)
    _abi.call("dwb_embed_decode", _ptr(seq), seq.stride(0), _ptr(pos_dev), _ptr(E), _ptr(P), int(E.dtype == F32), _ptr(x), B, d, vocab,
              _stream())
    return x


def attention_decode(q, k_new, v_new, k_cache, v_cache, cache_rows, B, H, *, fixed_len=0, pos_dev=None):
    """One query row per (batch, head) against the cache.  k_cache / v_cache: bf16 [B*cache_rows, >=H*64] views with the same
    row pitch.  Self-attention: k_new / v_new ([B, .] views) are appended at row pos; cross-attention: fixed_len rows."""
    for t, n in ((q, "q"), (k_cache, "k_cache"), (v_cache, "v_cache")):
        _check2d(t, BF16, "attention_decode " + n)
    if k_cache.stride(0) != v_cache.stride(0):
        raise ValueError("attention_decode: K and V cache views must share a row pitch")
    if k_new is not None and k_new.stride(0) != v_new.stride(0):
        raise ValueError("attention_decode: k_new and v_new must share a row pitch")
    out = torch.empty((B, H * HEAD_DIM), dtype=BF16, device=q.device)
    _abi.call("dwb_attention_decode", _ptr(q), q.stride(0), _ptr(k_new), _ptr(v_new), 0 if k_new is None else k_new.stride(0), _ptr(k_cache),
              _ptr(v_cache), k_cache.stride(0), int(cache_rows), _ptr(out), out.stride(0), B, H, HEAD_DIM, int(fixed_len), _ptr(pos_dev),
              HEAD_DIM ** -0.5, _stream())
    return out


def greedy_pick(logits, vocab, bias_all, bias_begin, begin_pos, seq, prompt_len, finished, eos, pad, pos_dev):
    _abi.call("dwb_greedy_pick", _ptr(logits), logits.stride(0), vocab, _ptr(bias_all), _ptr(bias_begin), int(begin_pos), _ptr(seq),
              seq.stride(0), int(prompt_len), _ptr(finished), int(eos), int(pad), _ptr(pos_dev), seq.shape[0], _stream())


def greedy_pick_timestamps(logits, vocab, bias_all, bias_begin, begin_pos, seq, prompt_len, finished, eos, pad, pos_dev, timestamp_begin,
                           max_initial_timestamp_index):
    _abi.call("dwb_greedy_pick_timestamps", _ptr(logits), logits.stride(0), vocab, _ptr(bias_all), _ptr(bias_begin), int(begin_pos), _ptr(seq),
              seq.stride(0), int(prompt_len), _ptr(finished), int(eos), int(pad), _ptr(pos_dev), seq.shape[0], int(timestamp_begin),
              -1 if max_initial_timestamp_index is None else int(max_initial_timestamp_index), _stream())


def decode_advance(pos_dev, finished, done_at):
    _abi.call("dwb_decode_advance", _ptr(pos_dev), _ptr(finished), finished.numel(), _ptr(done_at), _stream())


def collate_labels(tokens, lengths, decoder_start_token_id):
    """tokens int64 [B, L+1] (padded), lengths int32 [B] (CUDA) -> (decoder_input_ids, labels) int64 [B, L]."""
    assert tokens.dtype == torch.int64 and tokens.is_contiguous() and lengths.dtype == torch.int32 and tokens.is_cuda and lengths.is_cuda
    B, L1 = tokens.shape
    dec_in = torch.empty((B, L1 - 1), dtype=torch.int64, device=tokens.device)
    labels = torch.empty((B, L1 - 1), dtype=torch.int64, device=tokens.device)
    _abi.call("dwb_collate_labels", _ptr(tokens), _ptr(lengths), B, L1, int(decoder_start_token_id), _ptr(dec_in), _ptr(labels), _stream())
    return dec_in, labels


def colsum(m, out=None, accumulate=False):
    _check2d(m, BF16, "colsum")
    if out is None:
        out = torch.empty((m.shape[1],), dtype=F32, device=m.device)
    _abi.call("dwb_colsum_bf16", _ptr(m), m.stride(0), _ptr(out), m.shape[0], m.shape[1], int(accumulate), _stream())
    return out


def gelu_bwd(da, h):
    assert da.is_contiguous() and h.is_contiguous()
    dh = torch.empty_like(da)
    _abi.call("dwb_gelu_bwd", _ptr(da), _ptr(h), _ptr(dh), da.numel(), _stream())
    return dh


def gelu_fwd(h):
    y = torch.empty_like(h)
    _abi.call("dwb_gelu_fwd", _ptr(h), _ptr(y), h.numel(), _stream())
    return y


def kd_loss(student_logits, teacher_logits, labels, vocab, temperature, ce_weight, kl_weight, want_grad=True):
    """student_logits/teacher_logits: fp32 [rows, ld>=vocab] buffers.  Returns (metrics4 device tensor, dlogits bf16 or None)."""
    rows, ld = student_logits.shape
    dev = student_logits.device
    ws = torch.empty((int(_abi.call("dwb_kd_loss_workspace_bytes", rows)),), dtype=torch.uint8, device=dev)
    metrics = torch.empty((4,), dtype=F32, device=dev)
    dl = None
    ldd = round_up(vocab, 8)
    if want_grad:
        dl = torch.empty((rows, ldd), dtype=BF16, device=dev)
    labels = labels.reshape(-1).contiguous()
    _abi.call("dwb_kd_loss", _ptr(student_logits), _ptr(teacher_logits), student_logits.stride(0), _ptr(labels), rows, vocab,
              float(temperature), float(ce_weight), float(kl_weight), _ptr(metrics), _ptr(dl), ldd, _ptr(ws), _stream())
    return metrics, dl


def scale_bf16_dev(x, scale_dev):
    """x (contiguous bf16) *= scale_dev (0-d / 1-element fp32 CUDA tensor); free when the scalar is 1."""
    if not (x.is_contiguous() and x.dtype == BF16 and x.numel() % 8 == 0):
        raise ValueError("scale_bf16_dev: contiguous bf16 buffer with numel % 8 == 0 required")
    if scale_dev.dtype != F32 or scale_dev.numel() != 1 or not scale_dev.is_cuda:
        raise ValueError("scale_bf16_dev: the scale must be a one-element fp32 CUDA tensor")
    _abi.call("dwb_scale_bf16_dev", _ptr(x), x.numel(), _ptr(scale_dev), _stream())
    return x


def grad_sumsq(g, out):
    _abi.call("dwb_grad_sumsq", _ptr(g), g.numel(), _ptr(out), _stream())


def adamw_step(p, g, m, v, p_bf16, lr, beta1, beta2, eps, weight_decay, step, grad_sumsq_t, max_grad_norm, grad_scale=1.0,
               zero_grad=True):
    _abi.call("dwb_adamw_step", _ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(p_bf16), p.numel(), float(lr), float(beta1), float(beta2),
              float(eps), float(weight_decay), int(step), _ptr(grad_sumsq_t), float(max_grad_norm), float(grad_scale),
              int(zero_grad), _stream())
