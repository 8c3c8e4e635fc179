"""Per-kernel parity of the non-GEMM C-ABI entry points against plain PyTorch fp32 on the GPU."""
import math

import numpy as np

import pytest
import torch
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


@pytest.fixture(scope="module")
def ops():
    from distil_whisper_b200 import ops as o, _abi
    _abi.call("dwb_check_device")
    return o


def _randn(shape, seed, scale=1.0, dtype=torch.float32):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(dtype)


def _rel(x, y):
    return float((x.float() - y.float()).norm() / (y.float().norm() + 1e-20))


# ---------------------------------------------------------------------------------------------- attention
def _sdpa_ref(q, k, v, B, H, Sq, Sk, causal):
    qf = q.float().view(B, Sq, H, 64).transpose(1, 2)
    kf = k.float().view(B, Sk, H, 64).transpose(1, 2)
    vf = v.float().view(B, Sk, H, 64).transpose(1, 2)
    s = (qf @ kf.transpose(-1, -2)) * 0.125
    if causal:
        s = s.masked_fill(~torch.ones(Sq, Sk, dtype=torch.bool, device=q.device).tril(), float("-inf"))
    lse = torch.logsumexp(s, dim=-1)
    o = torch.softmax(s, dim=-1) @ vf
    return o.transpose(1, 2).reshape(B * Sq, H * 64), lse


@pytest.mark.parametrize("B,H,Sq,Sk,causal", [(2, 2, 128, 128, True), (1, 3, 12, 12, True), (2, 2, 100, 1500, False),
                                              (1, 2, 1500, 1500, False), (3, 1, 77, 50, False), (2, 20, 128, 128, True)])
def test_attention_fwd(ops, B, H, Sq, Sk, causal):
    d = H * 64
    qkv = _randn((B * Sq, 3 * d), 1, 1.0, torch.bfloat16)          # fused layout: strided views
    kv = _randn((B * Sk, 2 * d), 2, 1.0, torch.bfloat16)
    q = qkv[:, :d]
    k, v = (qkv[:, d:2 * d], qkv[:, 2 * d:]) if Sq == Sk else (kv[:, :d], kv[:, d:])
    o, lse = ops.attention_fwd(q, k, v, B, H, Sq, Sk, causal)
    o_ref, lse_ref = _sdpa_ref(q, k, v, B, H, Sq, Sk, causal)
    assert _rel(o, o_ref) < 8e-3, _rel(o, o_ref)
    assert torch.allclose(lse, lse_ref, atol=2e-3, rtol=1e-4)


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("B,H,Sq,Sk", [(2, 2, 1500, 1500), (1, 3, 300, 300), (2, 2, 128, 128), (1, 1, 100, 700), (3, 20, 50, 50)])
def test_attention_fwd_tcgen05(ops, B, H, Sq, Sk, causal):
    """tcgen05 / TMEM encoder attention vs torch fp32 and vs the mma.sync kernel (same C-ABI contract)."""
    d = H * 64
    qkv = _randn((B * Sq, 3 * d), 11, 1.0, torch.bfloat16)
    kv = _randn((B * Sk, 2 * d), 12, 1.0, torch.bfloat16)
    q = qkv[:, :d]
    k, v = (qkv[:, d:2 * d], qkv[:, 2 * d:]) if Sq == Sk else (kv[:, :d], kv[:, d:])
    o, lse = ops.attention_fwd(q, k, v, B, H, Sq, Sk, causal, use_tc=True)
    o_ref, lse_ref = _sdpa_ref(q, k, v, B, H, Sq, Sk, causal)
    assert _rel(o, o_ref) < 8e-3, _rel(o, o_ref)
    assert torch.allclose(lse, lse_ref, atol=2e-3, rtol=1e-4)
    o2, _ = ops.attention_fwd(q, k, v, B, H, Sq, Sk, causal, use_tc=False)
    assert _rel(o, o2.float()) < 8e-3


@pytest.mark.parametrize("growth", [3.0, -3.0, 0.0])
def test_attention_fwd_tcgen05_reference_maximum_jumps(ops, growth):
    """Scores whose row maximum climbs by ~2^30 from one 128-key tile to the next (growth > 0: every tile takes the kernel's
    slow path -- row sum above 2^15 -> true maximum, O / l rescaled, exponentials redone), falls (growth < 0: later tiles
    underflow against the first tile's reference) or sits at huge constant magnitude (0)."""
    B, H, S = 2, 2, 700
    d = H * 64
    q = _randn((B * S, d), 21, 3.0, torch.bfloat16)
    k = _randn((B * S, d), 22, 3.0, torch.bfloat16).float().view(B, S, d)
    tile = (torch.arange(S, device="cuda") // 128).float()
    if growth > 0:
        k = k * (1.0 + tile)[None, :, None]
    elif growth < 0:
        k = k * (6.0 - tile)[None, :, None]
    else:
        k = k * 4.0
    k = k.reshape(B * S, d).to(torch.bfloat16)
    v = _randn((B * S, d), 23, 1.0, torch.bfloat16)
    o, lse = ops.attention_fwd(q, k, v, B, H, S, S, False, use_tc=True)
    o_ref, lse_ref = _sdpa_ref(q, k, v, B, H, S, S, False)
    assert torch.isfinite(o.float()).all() and torch.isfinite(lse).all()
    assert _rel(o, o_ref) < 1e-2, _rel(o, o_ref)
    assert torch.allclose(lse, lse_ref, atol=5e-2, rtol=2e-4)


@pytest.mark.parametrize("use_tc", [True, False])
@pytest.mark.parametrize("B,H,Sq,Sk,causal", [(2, 2, 128, 128, True), (1, 2, 12, 12, True), (2, 2, 100, 300, False),
                                              (1, 1, 200, 1500, False), (2, 3, 70, 70, True), (1, 2, 1500, 1500, False),
                                              (2, 1, 300, 300, True), (1, 2, 50, 260, True)])
def test_attention_bwd(ops, B, H, Sq, Sk, causal, use_tc):
    """tcgen05 backward (use_tc) and the mma.sync backward against torch autograd of the fp32 reference."""
    d = H * 64
    qkv = _randn((B * Sq, 3 * d), 3, 1.0, torch.bfloat16)            # strided views, like the fused QKV buffer
    q = qkv[:, d:2 * d]
    k = _randn((B * Sk, d), 4, 1.0, torch.bfloat16)
    v = _randn((B * Sk, d), 5, 1.0, torch.bfloat16)
    dout = _randn((B * Sq, d), 6, 1.0, torch.bfloat16)
    o, lse = ops.attention_fwd(q, k, v, B, H, Sq, Sk, causal)
    dq = torch.empty((B * Sq, d), dtype=torch.bfloat16, device="cuda")
    dkv = torch.empty((B * Sk, 2 * d), dtype=torch.bfloat16, device="cuda")
    dk, dv = dkv[:, :d], dkv[:, d:]
    ops.attention_bwd(q, k, v, o, dout, lse, B, H, Sq, Sk, causal, dq, dk, dv, use_tc=use_tc)
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    o_ref, _ = _sdpa_ref(qf, kf, vf, B, H, Sq, Sk, causal)
    o_ref.backward(dout.float())
    assert _rel(dq, qf.grad) < 1.5e-2, ("dq", _rel(dq, qf.grad))
    assert _rel(dk, kf.grad) < 1.5e-2, ("dk", _rel(dk, kf.grad))
    assert _rel(dv, vf.grad) < 1.5e-2, ("dv", _rel(dv, vf.grad))


# ---------------------------------------------------------------------------------------------- layernorm
@pytest.mark.parametrize("rows,d,mod", [(37, 128, 0), (1000, 1280, 0), (300, 768, 100), (64, 1024, 0)])
def test_add_layernorm(ops, rows, d, mod):
    x = _randn((mod if mod else rows, d), 1, 2.0)
    y = _randn((rows, d), 2, 1.0, torch.bfloat16)
    g, b = _randn((d,), 3) * 0.1 + 1, _randn((d,), 4) * 0.1
    x_new, ln, mean, rstd = ops.add_layernorm(x, y, g, b, rows=rows, d=d, x_rows_mod=mod, save_stats=True)
    xin = x.repeat(rows // mod, 1) if mod else x
    ref_x = xin + y.float()
    assert torch.allclose(x_new, ref_x, atol=1e-6)
    ref_ln = F.layer_norm(ref_x, (d,), g, b, 1e-5)
    assert _rel(ln, ref_ln) < 4e-3
    assert torch.allclose(mean, ref_x.mean(-1), atol=1e-5)
    assert torch.allclose(rstd, (ref_x.var(-1, unbiased=False) + 1e-5).rsqrt(), rtol=1e-4)
    # no-add variant
    _, ln2, _, _ = ops.add_layernorm(xin.contiguous(), None, g, b, rows=rows, d=d, write_x=False)
    assert _rel(ln2, F.layer_norm(xin, (d,), g, b, 1e-5)) < 4e-3


@pytest.mark.parametrize("rows,d", [(50, 128), (4096, 1280), (333, 768)])
def test_layernorm_bwd(ops, rows, d):
    x = _randn((rows, d), 1, 2.0)
    dy = _randn((rows, d), 2, 1.0, torch.bfloat16)
    dres = _randn((rows, d), 5, 1.0)
    g, b = _randn((d,), 3) * 0.1 + 1, _randn((d,), 4) * 0.1
    _, _, mean, rstd = ops.add_layernorm(x, None, g, b, rows=rows, d=d, write_x=False, save_stats=True)
    dgamma = torch.zeros(d, device="cuda")
    dbeta = torch.zeros(d, device="cuda")
    dx, dxb = ops.layernorm_bwd(dy, x, mean, rstd, g, dres, dgamma, dbeta, rows=rows, d=d)
    xr, gr, br = x.clone().requires_grad_(True), g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    F.layer_norm(xr, (d,), gr, br, 1e-5).backward(dy.float())
    assert _rel(dx, xr.grad + dres) < 1e-4, _rel(dx, xr.grad + dres)
    assert _rel(dxb, xr.grad + dres) < 5e-3
    assert _rel(dgamma, gr.grad) < 1e-4 and _rel(dbeta, br.grad) < 1e-4


# ---------------------------------------------------------------------------------------------- conv stem
@pytest.mark.parametrize("B,C,L,d", [(2, 80, 100, 128), (1, 128, 3000, 256), (3, 80, 64, 64)])
def test_conv_stem_as_gemm(ops, B, C, L, d):
    mel = _randn((B, C, L), 1)
    w1, b1 = _randn((d, C, 3), 2, 0.05), _randn((d,), 3, 0.05)
    w2, b2 = _randn((d, d, 3), 4, 0.05), _randn((d,), 5, 0.05)
    a1 = ops.im2col_conv1(mel)
    w1b = torch.zeros((d, a1.shape[1]), dtype=torch.bfloat16, device="cuda")
    ops.cast_f32_to_bf16(w1.reshape(d, C * 3), w1b[:, :C * 3])
    x1 = ops.gemm(a1, w1b, bias=b1, act=1)                                   # [B*L, d] channels-last
    ref1 = F.gelu(F.conv1d(mel.bfloat16().float(), w1.bfloat16().float(), b1, padding=1))   # [B, d, L]
    assert _rel(x1.view(B, L, d).transpose(1, 2), ref1) < 8e-3
    a2 = ops.im2col_conv2(x1, B, L, d)
    w2b = ops.conv_weight_to_kc(w2)
    x2 = ops.gemm(a2, w2b, bias=b2, act=1)                                   # [B*L/2, d]
    ref2 = F.gelu(F.conv1d(x1.float().view(B, L, d).transpose(1, 2), w2.bfloat16().float(), b2, stride=2, padding=1))
    assert _rel(x2.view(B, L // 2, d).transpose(1, 2), ref2) < 8e-3
    # weight-gradient layout round trip
    gk = _randn((d, 3 * d), 6)
    dw = ops.conv_wgrad_kc_to_ck(gk, d, d)
    assert torch.equal(dw, gk.view(d, 3, d).permute(0, 2, 1).contiguous())


def test_col2im_conv2_gelu_bwd(ops):
    """Input gradient of the stride-2 conv (col2im over overlapping windows) fused with conv1's GELU backward."""
    B, L, d = 2, 20, 64
    pre1 = _randn((B * L, d), 1, 1.5, torch.bfloat16)
    g = _randn((B * (L // 2), 3 * d), 2, 1.0, torch.bfloat16)
    out = ops.col2im_conv2_gelu_bwd(g, pre1, B, L, d)
    pr = pre1.float().view(B, L, d).requires_grad_(True)
    x = F.gelu(pr)
    xp = F.pad(x, (0, 0, 1, 0))                                      # position -1 -> zeros
    cols = torch.stack([xp[:, k:k + L:2, :][:, :L // 2] for k in range(3)], dim=2)   # [B, L/2, 3, d]: taps 2t-1+k
    (cols.reshape(B * (L // 2), 3 * d) * g.float()).sum().backward()
    assert _rel(out, pr.grad.view(B * L, d)) < 6e-3, _rel(out, pr.grad.view(B * L, d))
    # and the forward im2col agrees with the same tap construction
    a2 = ops.im2col_conv2(x.detach().bfloat16().reshape(B * L, d).contiguous(), B, L, d)
    assert torch.equal(a2, cols.detach().bfloat16().reshape(B * (L // 2), 3 * d))


# ---------------------------------------------------------------------------------------------- embeddings / small ops
@pytest.mark.parametrize("table_dtype", [torch.float32, torch.bfloat16])
def test_embedding_fwd_bwd(ops, table_dtype):
    B, T, d, V, pad = 3, 12, 128, 515, 500
    E = _randn((V, d), 1, 0.02).to(table_dtype)
    P = _randn((32, d), 2, 0.02).to(table_dtype)
    ids = torch.randint(0, V, (B, T), device="cuda")
    ids[0, -3:] = pad
    x = ops.embed_fwd(ids.contiguous(), E, P, B, T, d, V)
    ref = E.float()[ids] + P.float()[:T]
    assert torch.allclose(x.view(B, T, d), ref, atol=1e-6)
    dx = _randn((B * T, d), 3)
    dE, dP = torch.zeros((V, d), device="cuda"), torch.zeros((32, d), device="cuda")
    ops.embed_bwd(ids.contiguous(), dx, dE, dP, B, T, d, V, pad)
    Er = E.float().clone().requires_grad_(True)
    Pr = P.float().clone().requires_grad_(True)
    (F.embedding(ids, Er, padding_idx=pad) + Pr[:T]).backward(dx.view(B, T, d))
    assert torch.allclose(dE, Er.grad, atol=1e-5) and torch.allclose(dP, Pr.grad, atol=1e-5)
    assert float(dE[pad].abs().sum()) == 0.0


def test_colsum_gelu_cast(ops):
    m = _randn((1000, 520), 1, 1.0, torch.bfloat16)
    view = m[:, :514]
    out = ops.colsum(view)
    assert torch.allclose(out, view.float().sum(0), atol=2e-3, rtol=1e-4)
    ops.colsum(view, out=out, accumulate=True)
    assert torch.allclose(out, 2 * view.float().sum(0), atol=4e-3, rtol=1e-4)
    h = _randn((64, 256), 2, 2.0, torch.bfloat16)
    da = _randn((64, 256), 3, 1.0, torch.bfloat16)
    hr = h.float().requires_grad_(True)
    F.gelu(hr).backward(da.float())
    assert _rel(ops.gelu_bwd(da, h), hr.grad) < 6e-3
    assert _rel(ops.gelu_fwd(h), F.gelu(h.float())) < 6e-3
    src = _randn((33, 64), 4)
    dst = torch.zeros((33, 200), dtype=torch.bfloat16, device="cuda")
    ops.cast_f32_to_bf16(src, dst[:, 64:128], scale=0.5)
    assert torch.equal(dst[:, 64:128], (src * 0.5).bfloat16()) and float(dst[:, :64].abs().sum()) == 0
    back = ops.cast_bf16_to_f32(dst[:, 64:128])
    assert torch.equal(back, dst[:, 64:128].float())


# ---------------------------------------------------------------------------------------------- KD loss
@pytest.mark.parametrize("rows,V,T", [(36, 515, 2.0), (64, 51866, 2.0), (20, 1000, 1.0), (16, 515, 3.5)])
def test_kd_loss_against_reference_formula(ops, rows, V, T):
    from oracle import whisper_oracle as wo
    ld = ops.round_up(V, 8)
    s_buf = _randn((rows, ld), 1, 3.0)
    t_buf = _randn((rows, ld), 2, 3.0)
    s_buf[:, V:] = float("nan")            # padding columns must never be read into the result
    t_buf[:, V:] = float("nan")
    labels = torch.randint(0, V, (rows,), device="cuda")
    labels[::5] = -100
    metrics, dl = ops.kd_loss(s_buf, t_buf, labels, V, T, 0.8, 1.0)
    s = s_buf[:, :V].clone().requires_grad_(True)
    t = t_buf[:, :V]
    ce = F.cross_entropy(s, labels)
    kl = wo.kl_divergence(F.softmax(t / T, -1), F.log_softmax(s / T, -1), labels) * T ** 2      # ref :1453-1462,:1486-1490
    loss = 0.8 * ce + 1.0 * kl
    loss.backward()
    m = metrics.cpu()
    assert math.isclose(m[1], ce.item(), rel_tol=2e-5) and math.isclose(m[2], kl.item(), rel_tol=2e-4, abs_tol=1e-7)
    assert math.isclose(m[0], loss.item(), rel_tol=2e-5) and int(m[3]) == int((labels >= 0).sum())
    assert _rel(dl[:, :V], s.grad) < 5e-3, _rel(dl[:, :V], s.grad)
    assert float(dl[:, V:].float().abs().sum()) == 0.0
    assert float(dl[::5].float().abs().sum()) == 0.0
    # CE-only (no teacher) == HF CrossEntropyLoss path
    m2, _ = ops.kd_loss(s_buf, None, labels, V, 1.0, 1.0, 0.0, want_grad=False)
    assert math.isclose(m2.cpu()[0], ce.item(), rel_tol=2e-5)


# ---------------------------------------------------------------------------------------------- optimiser
@pytest.mark.parametrize("wd,max_norm", [(0.0, 1.0), (0.01, 0.05), (0.0, 0.0)])
def test_adamw_matches_torch(ops, wd, max_norm):
    n = 100003
    p0, grads = _randn((n,), 1), [_randn((n,), 10 + i, 0.3) for i in range(3)]
    p_ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([p_ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    pb = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    for step, g in enumerate(grads, 1):
        p_ref.grad = g.clone()
        if max_norm > 0:
            torch.nn.utils.clip_grad_norm_([p_ref], max_norm)
        opt.step()
        gg = g.clone()
        ss = torch.zeros(1, device="cuda")
        ops.grad_sumsq(gg, ss)
        assert math.isclose(float(ss), float((g.double() ** 2).sum()), rel_tol=1e-4)
        ops.adamw_step(p, gg, m, v, pb, 1e-3, 0.9, 0.999, 1e-8, wd, step, ss if max_norm > 0 else None, max_norm)
        assert float(gg.abs().sum()) == 0.0
    assert torch.allclose(p, p_ref.detach(), atol=2e-6, rtol=1e-5)
    assert torch.equal(pb, p.bfloat16())


def test_decode_attention_and_greedy_pick_against_torch():
    """Single-token decode kernels (csrc/decode.cu): cache append + attention over pos+1 keys, cross-attention over a fixed
    length, and the token pick with suppress biases / prompt / EOS bookkeeping, against plain torch."""
    from distil_whisper_b200 import ops
    torch.manual_seed(0)
    B, H, d, Tmax, S = 5, 3, 192, 40, 1500
    pos = 17
    pos_dev = torch.tensor([pos], dtype=torch.int32, device="cuda")
    qkv = (torch.randn(B, 3 * d, device="cuda") * 0.7).bfloat16()
    cache = (torch.randn(B * Tmax, 2 * d, device="cuda") * 0.7).bfloat16()
    ref_cache = cache.clone().view(B, Tmax, 2 * d)
    o = ops.attention_decode(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], cache[:, :d], cache[:, d:], Tmax, B, H, pos_dev=pos_dev)
    ref_cache[:, pos, :d] = qkv[:, d:2 * d]
    ref_cache[:, pos, d:] = qkv[:, 2 * d:]
    assert torch.equal(cache.view(B, Tmax, 2 * d)[:, :pos + 1], ref_cache[:, :pos + 1])           # appended in place, nothing else touched
    assert torch.equal(cache.view(B, Tmax, 2 * d)[:, pos + 1:], ref_cache[:, pos + 1:])

    def ref_attn(q, k, v):      # q [B, d], k/v [B, L, d]
        qh = q.float().view(B, H, 1, 64)
        kh = k.float().view(B, -1, H, 64).transpose(1, 2)
        vh = v.float().view(B, -1, H, 64).transpose(1, 2)
        p = torch.softmax(qh @ kh.transpose(-1, -2) / 8.0, dim=-1)
        return (p @ vh).transpose(1, 2).reshape(B, H * 64)
    want = ref_attn(qkv[:, :d], ref_cache[:, :pos + 1, :d], ref_cache[:, :pos + 1, d:])
    assert (o.float() - want).abs().max() < 2e-2 * want.abs().max()
    ckv = (torch.randn(B * S, 2 * d, device="cuda") * 0.7).bfloat16()
    o2 = ops.attention_decode(qkv[:, :d], None, None, ckv[:, :d], ckv[:, d:], S, B, H, fixed_len=S)
    want2 = ref_attn(qkv[:, :d], ckv.view(B, S, 2 * d)[:, :, :d], ckv.view(B, S, 2 * d)[:, :, d:])
    assert (o2.float() - want2).abs().max() < 2e-2 * want2.abs().max()
    # ---- greedy pick
    V, ld = 1003, 1008
    logits = torch.randn(B, ld, device="cuda")
    logits[:, V:] = 1e9                                   # padding columns must never win
    bias_all = torch.zeros(V, device="cuda")
    bias_begin = torch.zeros(V, device="cuda")
    top = logits[:, :V].argmax(-1)
    bias_all[top[0]] = float("-inf")                      # row 0's winner is suppressed at every step
    bias_begin[top[1]] = float("-inf")                    # row 1's winner only at the first generated position
    seq = torch.full((B, 8), 7, dtype=torch.int64, device="cuda")
    finished = torch.zeros(B, dtype=torch.int32, device="cuda")
    finished[2] = 1
    done = torch.zeros(1, dtype=torch.int32, device="cuda")
    eos = int(top[3])
    p = torch.tensor([2], dtype=torch.int32, device="cuda")           # consumed column 2 -> writes column 3; prompt_len 3 = begin_pos
    ops.greedy_pick(logits, V, bias_all, bias_begin, 3, seq, 3, finished, eos, 99, p)
    ops.decode_advance(p, finished, done)
    l2 = logits[:, :V].clone()
    l2[:, top[0]] = float("-inf")
    l2b = l2.clone()
    l2b[:, top[1]] = float("-inf")
    want = l2b.argmax(-1)
    assert seq[0, 3] == want[0] and seq[1, 3] == want[1] and seq[2, 3] == 99 and seq[3, 3] == eos and seq[4, 3] == want[4]
    assert finished.tolist() == [0, 0, 1, 1, 0] and int(p) == 3 and int(done) == 0 and (seq[:, :3] == 7).all()
    ops.greedy_pick(logits, V, bias_all, bias_begin, 3, seq, 3, finished, eos, 99, p)      # next position: begin bias no longer applies
    assert seq[1, 4] == l2.argmax(-1)[1] and seq[3, 4] == 99
    # inside the prompt nothing is written
    p0 = torch.tensor([0], dtype=torch.int32, device="cuda")
    before = seq.clone()
    ops.greedy_pick(logits, V, bias_all, bias_begin, 3, seq, 3, finished, eos, 99, p0)
    assert torch.equal(seq, before)
    finished.fill_(1)
    ops.decode_advance(p, finished, done)
    assert int(done) == int(p) + 1


def test_device_collator_matches_reference_collator_semantics():
    """dwb_collate_labels + DataCollatorSpeechSeq2SeqWithPaddingB200 against the oracle's restatement of the reference collator
    (ref:training/run_distillation.py:438-478): shift, -100 on padding, -100 on the prompt up to and including SOT; raw audio
    in -> log-mel computed on the device equals the extractor's own output."""
    from distil_whisper_b200.data import DataCollatorSpeechSeq2SeqWithPaddingB200
    from distil_whisper_b200.feature_extraction import WhisperFeatureExtractorB200
    from oracle import logmel_oracle as lo
    from oracle import whisper_oracle as wo
    sot, prev, pad = 501, 502, 500
    rows = [[sot, 7, 8, 9, 10],                              # plain row
            [prev, 3, 4, sot, 11, 12, 13],                   # prompted row: -100 up to and including SOT
            [sot, 5],                                        # short row
            [prev, 1, sot, 2, sot, 3, 4, 5, 6, 7, 8, 9]]     # a second SOT later in the row: only the first counts
    wav = lo.synthetic_waveforms(len(rows), seed=3, ragged=True)
    fe = WhisperFeatureExtractorB200(80)
    coll = DataCollatorSpeechSeq2SeqWithPaddingB200(None, sot, prev, max_target_length=14, pad_token_id=pad, feature_extractor=fe)
    feats_in = [{"input_values": wav[i][: 480000 - 1000 * i], "labels": r} for i, r in enumerate(rows)]
    batch = coll(feats_in)
    dec_ref, lab_ref = wo.collate_labels(rows, pad, sot, max_len=14)
    assert batch["decoder_input_ids"].is_cuda and torch.equal(batch["decoder_input_ids"].cpu(), dec_ref)
    assert torch.equal(batch["labels"].cpu(), lab_ref), (batch["labels"].cpu(), lab_ref)
    # the fixture exercises the prompt mask: row 1's labels are [3, 4, SOT, 11, 12, 13] -> masked up to and including SOT
    assert (lab_ref[1, :3] == -100).all() and lab_ref[1, 3] == 11 and (lab_ref[3, :2] == -100).all() and lab_ref[3, 2] == 2
    want = fe.extract_device(torch.from_numpy(fe.pad_or_trim([f["input_values"] for f in feats_in])).cuda())
    assert batch["input_features"].shape == (4, 80, 3000) and torch.equal(batch["input_features"], want)
    ref0 = lo.log_mel(fe.pad_or_trim([feats_in[1]["input_values"]]), 80)
    assert np.abs(batch["input_features"][1].cpu().numpy() - ref0[0]).max() < 2e-4
    # precomputed features (the reference's dataset layout) pass through unchanged
    b2 = coll([{"input_features": want[i].cpu().numpy(), "labels": r} for i, r in enumerate(rows)])
    assert torch.equal(b2["input_features"], want) and torch.equal(b2["labels"].cpu(), lab_ref)
    # longest padding
    coll2 = DataCollatorSpeechSeq2SeqWithPaddingB200(None, sot, prev, target_padding="longest", pad_token_id=pad, feature_extractor=fe)
    b3 = coll2([{"input_features": want[i], "labels": r} for i, r in enumerate(rows)])
    d3, l3 = wo.collate_labels(rows, pad, sot)
    assert torch.equal(b3["labels"].cpu(), l3) and torch.equal(b3["decoder_input_ids"].cpu(), d3)


def test_greedy_pick_with_timestamp_rules_matches_hf_processor_semantics():
    """dwb_greedy_pick_timestamps against oracle.timestamp_rules (== HF's WhisperTimeStampLogitsProcessor, pinned on CPU by
    tests/test_oracle.py) + suppress biases, on crafted prefixes covering every rule."""
    from distil_whisper_b200 import ops
    from oracle import whisper_oracle as wo
    V, ld, ts_begin, eos, begin = 140, 144, 130, 120, 3
    for max_init in (None, 1):
        for ids, logits in wo.synthetic_timestamp_cases(6, V, ts_begin, eos, begin, seed=3):
            B, t = ids.shape
            suppress, begin_suppress = [2, 11, 131], [eos, 7]
            sc = logits.clone()
            sc[:, suppress] = -float("inf")
            if t == begin:
                sc[:, begin_suppress] = -float("inf")
            want = wo.timestamp_rules(ids, sc, begin, ts_begin, eos, max_init).argmax(-1)
            buf = torch.full((B, ld), 1e9, device="cuda")
            buf[:, :V] = logits.cuda()
            bias_all = torch.zeros(V, device="cuda")
            bias_all[suppress] = float("-inf")
            bias_begin = torch.zeros(V, device="cuda")
            bias_begin[begin_suppress] = float("-inf")
            seq = torch.full((B, t + 2), 99, dtype=torch.int64, device="cuda")
            seq[:, :t] = ids.cuda()
            finished = torch.zeros(B, dtype=torch.int32, device="cuda")
            pos = torch.tensor([t - 1], dtype=torch.int32, device="cuda")
            ops.greedy_pick_timestamps(buf, V, bias_all, bias_begin, begin, seq, begin, finished, eos, 99, pos, ts_begin, max_init)
            assert torch.equal(seq[:, t].cpu(), want), (max_init, ids[0].tolist(), seq[:, t].cpu(), want)
            assert torch.equal(finished.cpu().bool(), want == eos)


@pytest.mark.parametrize("M,N,K,act,f32", [(32, 1280, 1280, 0, False), (16, 3840, 1280, 0, False), (64, 5120, 1280, 1, False), (48, 1280, 5120, 0, False),
                                          (32, 264, 80, 1, True), (3, 1280, 1280, 0, False)])
def test_skinny_decode_gemm_against_torch(M, N, K, act, f32):
    """dwb_gemm_skinny_bf16 (decode-step projections, M = batch rows) vs torch fp32 on the same bf16 operands; M = 3 takes the
    tcgen05 fallback inside ops.gemm_small_m."""
    from distil_whisper_b200 import ops
    torch.manual_seed(M + N)
    x = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    b = torch.randn(N, device="cuda") * 0.1
    out = ops.gemm_small_m(x, w, bias=b, act=act, out_dtype=torch.float32 if f32 else torch.bfloat16)
    ref = x.float() @ w.float().t() + b
    if act:
        ref = F.gelu(ref)
    assert out.shape == (M, N)
    tol = 2e-3 if f32 else 1e-2
    assert (out.float() - ref).abs().max() < tol * max(1.0, float(ref.abs().max())), float((out.float() - ref).abs().max())


def test_reversed_row_walk_gives_identical_results():
    """dwb_set_row_walk(1): GEMM tile walk (single-CTA and CTA-pair kernels), LayerNorm rows and attention batches run from the last row
    block to the first (L2 reuse along the encoder's kernel chain) -- the results must be bit-identical to the ascending walk."""
    from distil_whisper_b200 import _abi, ops
    torch.manual_seed(3)
    a = (torch.randn(3000, 1280, device="cuda") * 0.3).bfloat16()
    w = (torch.randn(3840, 1280, device="cuda") * 0.03).bfloat16()
    bias = torch.randn(3840, device="cuda") * 0.1
    x = torch.randn(3000, 1280, device="cuda")
    g, b_ = torch.randn(1280, device="cuda"), torch.randn(1280, device="cuda")
    B, H, S = 2, 20, 1500
    qkv = (torch.randn(B * S, 3 * H * 64, device="cuda") * 0.5).bfloat16()
    acc0 = torch.randn(3000, 3840, device="cuda")

    def run():
        outs = [ops.gemm(a, w, bias=bias, act=1), ops.gemm(a, w, bias=bias, impl=2), ops.gemm(a, w, bias=bias, impl=3)]
        acc = acc0.clone()
        ops.gemm(a, w, bias=bias, out=acc, accumulate=True)
        outs.append(acc)
        outs.append(ops.add_layernorm(x, None, g, b_, rows=3000, d=1280, write_x=False)[1])
        o, lse = ops.attention_fwd(qkv[:, :1280], qkv[:, 1280:2560], qkv[:, 2560:], B, H, S, S, False, use_tc=True)
        outs += [o, lse]
        return outs
    ref = run()
    _abi.call("dwb_set_row_walk", 1)
    try:
        rev = run()
    finally:
        _abi.call("dwb_set_row_walk", 0)
    for i, (r, v) in enumerate(zip(ref, rev)):
        assert torch.equal(r, v), i
