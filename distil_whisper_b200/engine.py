"""Forward / backward schedules of the Whisper encoder and decoder on the sm_100a kernels (ops.py -> libdwb.so).

Every arithmetic step of HF:models/whisper/modeling_whisper.py:593-647 (encoder), :691-796 (decoder), :995-1100
(LM head + CE) and of their autograd backward is issued here as an explicit sequence of kernel launches on raw
buffers.  torch.autograd only sees two coarse Functions (ModelForwardFn, CrossEntropyFn) plus the fused KD step
in kd.py, so `loss.backward()` from ref:training/run_distillation.py:1609 keeps working.

Numerics policy (DESIGN.md "precision"): bf16 tensor-core operands, fp32 accumulation, fp32 residual stream, fp32
LayerNorm statistics, fp32 logits and loss; parameters stay in their own dtype (fp32 master for the student, bf16
for the teacher, ref:training/run_distillation.py:985-1004) with bf16 shadows refreshed when a parameter changes.
"""
from __future__ import annotations

import torch

from . import ops

BF16, F32 = torch.bfloat16, torch.float32


# =================================================================================================================
# tracing: NVTX ranges per phase of the step -- visible in nsys / ncu --nvtx timelines, free otherwise
class nvtx_range:
    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        self.on = torch.cuda.is_available()
        if self.on:
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if self.on:
            torch.cuda.nvtx.range_pop()
        return False


# =================================================================================================================
# parameter shadows
_PARAM_EPOCH = [0]


def bump_param_epoch():
    """Called by optimisers that update parameters through raw pointers (optim.FusedAdamW): trainable parameters'
    bf16 shadows are rebuilt on next use.  torch optimisers bump Tensor._version themselves."""
    _PARAM_EPOCH[0] += 1

class _Cache:
    """bf16 / fp32 views of module parameters, rebuilt only when the parameter version or storage changes."""

    def __init__(self):
        self.store = {}

    @staticmethod
    def _key(params):
        # the epoch only concerns fp32 masters that optim.FusedAdamW rewrites through raw pointers
        return tuple((p.data_ptr(), p._version, p.dtype, _PARAM_EPOCH[0] if (p.requires_grad and p.dtype == F32) else 0)
                     for p in params if p is not None)

    def get(self, name, params, build):
        key = self._key(params)
        hit = self.store.get(name)
        if hit is not None and hit[0] == key:
            return hit[1]
        with torch.no_grad():
            val = build(hit[1] if hit is not None else None)
        self.store[name] = (key, val)
        return val


def _to_bf16(dst, p):
    """dst (bf16 [rows, cols] view) <- p (fp32 or bf16 parameter, 2-D after flattening trailing dims)."""
    src = p.detach().reshape(p.shape[0], -1) if p.dim() > 1 else p.detach().reshape(1, -1)
    if p.dtype == F32:
        ops.cast_f32_to_bf16(src, dst)
    else:
        dst.copy_(src)          # same-dtype device memcpy
    return dst


def bf16_of(cache, name, p):
    if p.dtype == BF16:
        return p.detach().reshape(p.shape[0], -1)

    def build(old):
        out = old if old is not None else torch.empty((p.shape[0], p[0].numel()), dtype=BF16, device=p.device)
        return _to_bf16(out, p)
    return cache.get(name, [p], build)


def f32_of(cache, name, p):
    if p.dtype == F32:
        return p.detach()

    def build(old):
        out = old if old is not None else torch.empty(p.shape, dtype=F32, device=p.device)
        ops.cast_bf16_to_f32(p.detach().reshape(1, -1), out.reshape(1, -1))
        return out
    return cache.get(name, [p], build)


def fused_rows_bf16(cache, name, plist):
    """Row-concatenate weights [n_i, k] into one bf16 [sum n_i, k] operand (fused QKV / KV projections)."""
    def build(old):
        rows = sum(p.shape[0] for p in plist)
        out = old if old is not None else torch.empty((rows, plist[0].shape[1]), dtype=BF16, device=plist[0].device)
        r = 0
        for p in plist:
            _to_bf16(out[r:r + p.shape[0]], p)
            r += p.shape[0]
        return out
    return cache.get(name, plist, build)


def fused_bias_f32(cache, name, blist, d):
    """Concatenate biases (None -> zeros: Whisper's k_proj has no bias) into one fp32 vector."""
    def build(old):
        out = old if old is not None else torch.zeros((len(blist) * d,), dtype=F32, device=next(b for b in blist if b is not None).device)
        for i, b in enumerate(blist):
            if b is None:
                continue
            seg = out[i * d:(i + 1) * d]
            if b.dtype == F32:
                seg.copy_(b.detach())
            else:
                ops.cast_bf16_to_f32(b.detach().reshape(1, -1), seg.reshape(1, -1))
        return out
    return cache.get(name, [b for b in blist if b is not None], build)


class _State:
    def __init__(self, module):
        self.m = module
        self.cache = _Cache()


def state_of(module) -> _State:
    st = module.__dict__.get("_dwb_state")
    if st is None:
        st = _State(module)
        module.__dict__["_dwb_state"] = st
    return st


def invalidate(model):
    for m in model.modules():
        m.__dict__.pop("_dwb_state", None)


def shift_tokens_right(input_ids, pad_token_id, decoder_start_token_id):
    """Integer index shuffling only (HF:models/whisper/modeling_whisper.py:68-81): teacher decoder inputs from labels."""
    out = input_ids.new_zeros(input_ids.shape)
    out[:, 1:] = input_ids[:, :-1]
    out[:, 0] = decoder_start_token_id
    out.masked_fill_(out == -100, pad_token_id)
    return out


def _grad_buf(p):
    """fp32 gradient buffer of a parameter (allocated zeroed on first use; kernels accumulate into it)."""
    if p.grad is None:
        p.grad = torch.zeros(p.shape, dtype=F32, device=p.device)
    return p.grad


# =================================================================================================================
# encoder
def _attn_weights(st, key, attn, fuse_qkv):
    c, d = st.cache, attn.q_proj.weight.shape[0]
    out = {}
    if fuse_qkv:
        out["wqkv"] = fused_rows_bf16(c, key + ".wqkv", [attn.q_proj.weight, attn.k_proj.weight, attn.v_proj.weight])
        out["bqkv"] = fused_bias_f32(c, key + ".bqkv", [attn.q_proj.bias, None, attn.v_proj.bias], d)
    else:
        out["wq"] = bf16_of(c, key + ".wq", attn.q_proj.weight)
        out["bq"] = f32_of(c, key + ".bq", attn.q_proj.bias)
        out["wkv"] = fused_rows_bf16(c, key + ".wkv", [attn.k_proj.weight, attn.v_proj.weight])
        out["bkv"] = fused_bias_f32(c, key + ".bkv", [None, attn.v_proj.bias], d)
    out["wo"] = bf16_of(c, key + ".wo", attn.out_proj.weight)
    out["bo"] = f32_of(c, key + ".bo", attn.out_proj.bias)
    return out


def _ln(st, key, ln):
    return f32_of(st.cache, key + ".g", ln.weight), f32_of(st.cache, key + ".b", ln.bias)


USE_TC_ATTENTION = True     # tcgen05 encoder attention (attention_tcgen05.cu); False -> mma.sync kernel


class _RowWalk:
    """Alternates the row-walk direction of consecutive streaming kernels (dwb_set_row_walk): every tensor of the 32-utterance encoder
    pass (123 - 491 MB) is larger than L2, so a consumer that starts where its producer finished finds ~100 MB of its input still in
    L2 instead of none.  DWB_ROW_WALK=0 keeps every kernel ascending (A/B)."""

    def __init__(self, first_reverse=True):
        import os
        self.on = os.environ.get("DWB_ROW_WALK", "1") != "0"
        self.rev = first_reverse

    def step(self):
        if self.on:
            from . import _abi
            _abi.call("dwb_set_row_walk", int(self.rev))
            self.rev = not self.rev

    def reset(self):
        if self.on:
            from . import _abi
            _abi.call("dwb_set_row_walk", 0)


def _conv1_w(st, ld):
    enc = st.m
    d, C = enc.conv1.weight.shape[0], enc.conv1.weight.shape[1]

    def build(old):
        w = torch.zeros((d, ld), dtype=BF16, device=enc.conv1.weight.device) if old is None else old
        _to_bf16(w[:, :3 * C], enc.conv1.weight)
        return w
    return st.cache.get("conv1.w", [enc.conv1.weight], build)


def _conv2_w(st):
    enc = st.m

    def build(old):
        wf = enc.conv2.weight.detach()
        return ops.conv_weight_to_kc(wf if wf.dtype == F32 else wf.float().contiguous())
    return st.cache.get("conv2.w", [enc.conv2.weight], build)


def encoder_forward(st: _State, input_features, save=False):
    """[B, n_mels, 2*S] fp32 -> LayerNorm'd hidden states as bf16 [B*S, d].  save=False is the inference schedule used when
    the encoder is frozen (the reference recipe, ref:training/run_distillation.py:1023-1026 / README `--freeze_encoder`);
    save=True keeps what encoder_backward needs (variant A: trainable encoder)."""
    enc = st.m
    cfg = enc.config
    if save:
        return encoder_forward_train(st, input_features)
    expected = cfg.max_source_positions * 2
    if input_features.shape[-1] != expected:       # HF:models/whisper/modeling_whisper.py:613-617
        raise ValueError(f"Whisper expects the mel input features to be of length {expected}, but found "
                         f"{input_features.shape[-1]}. Make sure to pad the input mel features to {expected}.")
    B, C, L = input_features.shape
    d, S, H = cfg.d_model, cfg.max_source_positions, cfg.encoder_attention_heads
    c = st.cache
    mel = input_features.to(F32).contiguous()
    a1 = ops.im2col_conv1(mel)                                          # [B*L, ld1]
    w1 = _conv1_w(st, a1.shape[1])
    x1 = ops.gemm(a1, w1, bias=f32_of(c, "conv1.b", enc.conv1.bias), act=1)            # gelu(conv1) [B*L, d]
    del a1
    a2 = ops.im2col_conv2(x1, B, L, d)                                                # [B*S, 3d]
    del x1
    w2 = _conv2_w(st)
    y = ops.gemm(a2, w2, bias=f32_of(c, "conv2.b", enc.conv2.bias), act=1)             # gelu(conv2) [B*S, d]
    del a2
    M = B * S
    # Residual stream x (fp32, [M, d]) lives in one buffer.  Layer 0's LayerNorm kernel forms x = positions + gelu(conv2);
    # after that every sub-layer output is ADDED INTO x by its GEMM epilogue (TMA reduce-add of the fp32 accumulator + bias),
    # so the LayerNorm kernels only read x and write the bf16 normalised rows -- half the HBM traffic of add+LN, and the
    # sub-layer output is never rounded to bf16 before the residual add.
    pos = f32_of(c, "pos", enc.embed_positions.weight)[:S]
    x = torch.empty((M, d), dtype=F32, device=mel.device)
    walk = _RowWalk(first_reverse=True)         # the conv2 GEMM above finished on the last rows of y
    try:
        for i, layer in enumerate(enc.layers):
            k = f"l{i}"
            w = _attn_weights(st, k + ".sa", layer.self_attn, fuse_qkv=True)
            g, b_ = _ln(st, k + ".ln1", layer.self_attn_layer_norm)
            walk.step()
            if i == 0:
                _, h, _, _ = ops.add_layernorm(pos, y, g, b_, rows=M, d=d, x_rows_mod=S, x_out=x)
                del y
            else:
                _, h, _, _ = ops.add_layernorm(x, None, g, b_, rows=M, d=d, write_x=False)
            walk.step()
            qkv = ops.gemm(h, w["wqkv"], bias=w["bqkv"])
            walk.step()
            o, _ = ops.attention_fwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], B, H, S, S, causal=False, out=h,
                                     need_lse=False, use_tc=USE_TC_ATTENTION)
            del qkv
            walk.step()
            ops.gemm(o, w["wo"], bias=w["bo"], out=x, accumulate=True)                       # x += out_proj(attn)
            g, b_ = _ln(st, k + ".ln2", layer.final_layer_norm)
            walk.step()
            _, h, _, _ = ops.add_layernorm(x, None, g, b_, rows=M, d=d, write_x=False)
            walk.step()
            a = ops.gemm(h, bf16_of(c, k + ".fc1.w", layer.fc1.weight), bias=f32_of(c, k + ".fc1.b", layer.fc1.bias), act=1)
            walk.step()
            ops.gemm(a, bf16_of(c, k + ".fc2.w", layer.fc2.weight), bias=f32_of(c, k + ".fc2.b", layer.fc2.bias), out=x,
                     accumulate=True)                                                          # x += fc2(gelu(fc1))
            del a
        g, b_ = _ln(st, "ln_f", enc.layer_norm)
        walk.step()
        if len(enc.layers) == 0:
            _, out, _, _ = ops.add_layernorm(pos, y, g, b_, rows=M, d=d, x_rows_mod=S, write_x=False)
        else:
            _, out, _, _ = ops.add_layernorm(x, None, g, b_, rows=M, d=d, write_x=False)
    finally:
        walk.reset()
    return out, None


def encoder_forward_train(st: _State, input_features):
    """Training schedule of the encoder: same arithmetic as encoder_forward, but pre-activations, LayerNorm inputs /
    statistics, fused QKV, attention outputs and log-sum-exps are kept for encoder_backward."""
    enc = st.m
    cfg = enc.config
    expected = cfg.max_source_positions * 2
    if input_features.shape[-1] != expected:
        raise ValueError(f"Whisper expects the mel input features to be of length {expected}, but found "
                         f"{input_features.shape[-1]}. Make sure to pad the input mel features to {expected}.")
    B, C, L = input_features.shape
    d, S, H = cfg.d_model, cfg.max_source_positions, cfg.encoder_attention_heads
    c = st.cache
    M = B * S
    mel = input_features.to(F32).contiguous()
    a1 = ops.im2col_conv1(mel)
    w1 = _conv1_w(st, a1.shape[1])
    pre1 = ops.gemm(a1, w1, bias=f32_of(c, "conv1.b", enc.conv1.bias))                   # conv1 pre-activation [B*L, d]
    x1 = ops.gelu_fwd(pre1)
    a2 = ops.im2col_conv2(x1, B, L, d)
    del x1
    w2 = _conv2_w(st)
    pre2 = ops.gemm(a2, w2, bias=f32_of(c, "conv2.b", enc.conv2.bias))                   # conv2 pre-activation [M, d]
    y = ops.gelu_fwd(pre2)
    ctx = {"B": B, "C": C, "L": L, "a1": a1, "pre1": pre1, "a2": a2, "pre2": pre2, "layers": []}
    x = f32_of(c, "pos", enc.embed_positions.weight)[:S]
    x_mod = S
    for i, layer in enumerate(enc.layers):
        k = f"l{i}"
        w = _attn_weights(st, k + ".sa", layer.self_attn, fuse_qkv=True)
        g, b_ = _ln(st, k + ".ln1", layer.self_attn_layer_norm)
        xa, h1, mu1, rs1 = ops.add_layernorm(x, y, g, b_, rows=M, d=d, x_rows_mod=x_mod, save_stats=True)
        x_mod = 0
        qkv = ops.gemm(h1, w["wqkv"], bias=w["bqkv"])
        o, lse = ops.attention_fwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], B, H, S, S, causal=False, need_lse=True,
                                   use_tc=USE_TC_ATTENTION)
        y1 = ops.gemm(o, w["wo"], bias=w["bo"])
        g, b_ = _ln(st, k + ".ln2", layer.final_layer_norm)
        xb, h2, mu2, rs2 = ops.add_layernorm(xa, y1, g, b_, rows=M, d=d, save_stats=True)
        del y1
        hpre = ops.gemm(h2, bf16_of(c, k + ".fc1.w", layer.fc1.weight), bias=f32_of(c, k + ".fc1.b", layer.fc1.bias))
        a = ops.gelu_fwd(hpre)
        y = ops.gemm(a, bf16_of(c, k + ".fc2.w", layer.fc2.weight), bias=f32_of(c, k + ".fc2.b", layer.fc2.bias))
        ctx["layers"].append(dict(xa=xa, h1=h1, mu1=mu1, rs1=rs1, qkv=qkv, o=o, lse=lse, xb=xb, h2=h2, mu2=mu2, rs2=rs2,
                                  hpre=hpre, a=a))
        x = xb
    g, b_ = _ln(st, "ln_f", enc.layer_norm)
    xf, out, muf, rsf = ops.add_layernorm(x, y, g, b_, rows=M, d=d, x_rows_mod=x_mod, save_stats=True)
    ctx.update(xf=xf, muf=muf, rsf=rsf)
    return out, ctx


def encoder_backward(st: _State, ctx, denc_bf16):
    """denc_bf16: bf16 [B*S, d] gradient wrt the encoder's LayerNorm'd output.  Accumulates every trainable encoder
    parameter gradient into .grad (autograd of HF:models/whisper/modeling_whisper.py:593-647)."""
    enc = st.m
    cfg = enc.config
    d, S, H = cfg.d_model, cfg.max_source_positions, cfg.encoder_attention_heads
    B, C, L = ctx["B"], ctx["C"], ctx["L"]
    M = B * S
    c = st.cache
    g, _ = _ln(st, "ln_f", enc.layer_norm)
    dx, dxb = _ln_bwd(denc_bf16, ctx["xf"], ctx["muf"], ctx["rsf"], enc.layer_norm, g, None, M, d)
    for i in reversed(range(len(enc.layers))):
        layer, Lc, k = enc.layers[i], ctx["layers"][i], f"l{i}"
        da = _linear_bwd(dxb, Lc["a"], bf16_of(c, k + ".fc2.w", layer.fc2.weight), layer.fc2.weight, layer.fc2.bias)
        dhpre = ops.gelu_bwd(da, Lc["hpre"])
        del da
        dh2 = _linear_bwd(dhpre, Lc["h2"], bf16_of(c, k + ".fc1.w", layer.fc1.weight), layer.fc1.weight, layer.fc1.bias)
        del dhpre
        g, _ = _ln(st, k + ".ln2", layer.final_layer_norm)
        dx, dxb = _ln_bwd(dh2, Lc["xb"], Lc["mu2"], Lc["rs2"], layer.final_layer_norm, g, dx, M, d)
        sa = layer.self_attn
        ws = _attn_weights(st, k + ".sa", sa, fuse_qkv=True)
        do = _linear_bwd(dxb, Lc["o"], ws["wo"], sa.out_proj.weight, sa.out_proj.bias)
        qkv = Lc["qkv"]
        dqkv = torch.empty((M, 3 * d), dtype=BF16, device=qkv.device)
        ops.attention_bwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], Lc["o"], do, Lc["lse"], B, H, S, S, False,
                          dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:])
        _linear_bwd(dqkv[:, :d], Lc["h1"], None, sa.q_proj.weight, sa.q_proj.bias, need_dx=False)
        _linear_bwd(dqkv[:, d:2 * d], Lc["h1"], None, sa.k_proj.weight, None, need_dx=False)
        _linear_bwd(dqkv[:, 2 * d:], Lc["h1"], None, sa.v_proj.weight, sa.v_proj.bias, need_dx=False)
        dh1 = ops.gemm(dqkv, ws["wqkv"], b_mn=True)
        del dqkv
        g, _ = _ln(st, k + ".ln1", layer.self_attn_layer_norm)
        dx, dxb = _ln_bwd(dh1, Lc["xa"], Lc["mu1"], Lc["rs1"], layer.self_attn_layer_norm, g, dx, M, d)
        ctx["layers"][i] = None                       # release this layer's activations
    # conv stem.  dxb = d loss / d gelu(conv2) (the position table is frozen)
    d2 = ops.gelu_bwd(dxb, ctx["pre2"])                                                   # [M, d] wrt conv2 pre-activation
    w2p = enc.conv2.weight
    if w2p.requires_grad:
        gkc = ops.gemm(d2, ctx["a2"], a_mn=True, b_mn=True, out_dtype=F32)                # [d, 3d] in (k, c) column order
        _abi_call_conv_wgrad(gkc, _grad_buf(w2p), d, d)
    if enc.conv2.bias.requires_grad:
        ops.colsum(d2, out=_grad_buf(enc.conv2.bias), accumulate=True)
    w2 = _conv2_w(st)
    da2 = ops.gemm(d2, w2, b_mn=True)                                                      # [M, 3d]
    d1 = ops.col2im_conv2_gelu_bwd(da2, ctx["pre1"], B, L, d)                              # [B*L, d] wrt conv1 pre-activation
    del da2
    w1p = enc.conv1.weight
    if w1p.requires_grad:
        ops.gemm(d1, ctx["a1"][:, :3 * C], a_mn=True, b_mn=True, out=_grad_buf(w1p).reshape(d, 3 * C), accumulate=True)
    if enc.conv1.bias.requires_grad:
        ops.colsum(d1, out=_grad_buf(enc.conv1.bias), accumulate=True)


def _abi_call_conv_wgrad(gkc, dw, O, Cc):
    from . import _abi
    import ctypes as C
    _abi.call("dwb_conv_wgrad_kc_to_ck", C.c_void_p(gkc.data_ptr()), C.c_void_p(dw.data_ptr()), O, Cc, 1,
              C.c_void_p(torch.cuda.current_stream().cuda_stream))


# =================================================================================================================
# decoder
def decoder_forward(st: _State, ids, enc, B, S, save, cross_kv=None):
    """ids [B, T] int64, enc bf16 [B*S, d] -> (final LayerNorm output bf16 [B*T, d], ctx for the backward).
    cross_kv: optional dict filled with / read from the per-layer cross-attention K/V projections of `enc` (they do not
    depend on the decoder inputs, so greedy decoding projects the 1500 encoder positions once, not once per token)."""
    dec = st.m
    cfg = dec.config
    d, H, T = cfg.d_model, cfg.decoder_attention_heads, ids.shape[1]
    if T > cfg.max_target_positions:
        raise ValueError(f"decoder sequence length {T} exceeds max_target_positions {cfg.max_target_positions}")
    c = st.cache
    M = B * T
    ids = ids.contiguous()
    E, P = dec.embed_tokens.weight, dec.embed_positions.weight
    if E.dtype != P.dtype:
        raise ValueError("embed_tokens and embed_positions must share a dtype")
    x = ops.embed_fwd(ids, E.detach(), P.detach(), B, T, d, cfg.vocab_size)            # fp32 [M, d]
    y = None
    ctx = {"layers": [], "ids": ids, "B": B, "T": T, "S": S, "enc": enc} if save else None
    for i, layer in enumerate(dec.layers):
        k = f"l{i}"
        L = {}
        # --- causal self-attention
        ws = _attn_weights(st, k + ".sa", layer.self_attn, fuse_qkv=True)
        g, b_ = _ln(st, k + ".ln1", layer.self_attn_layer_norm)
        x1, h1, mu1, rs1 = ops.add_layernorm(x, y, g, b_, rows=M, d=d, save_stats=save)
        qkv = ops.gemm(h1, ws["wqkv"], bias=ws["bqkv"])
        o1, lse1 = ops.attention_fwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], B, H, T, T, causal=True, need_lse=save,
                                     use_tc=USE_TC_ATTENTION)
        y1 = ops.gemm(o1, ws["wo"], bias=ws["bo"])
        # --- cross-attention over the encoder states
        wc = _attn_weights(st, k + ".ca", layer.encoder_attn, fuse_qkv=False)
        g, b_ = _ln(st, k + ".ln2", layer.encoder_attn_layer_norm)
        x2, h2, mu2, rs2 = ops.add_layernorm(x1, y1, g, b_, rows=M, d=d, save_stats=save)
        qc = ops.gemm(h2, wc["wq"], bias=wc["bq"])
        if cross_kv is not None and i in cross_kv:
            kvc = cross_kv[i]
        else:
            kvc = ops.gemm(enc, wc["wkv"], bias=wc["bkv"])                              # [B*S, 2d]
            if cross_kv is not None:
                cross_kv[i] = kvc
        o2, lse2 = ops.attention_fwd(qc, kvc[:, :d], kvc[:, d:], B, H, T, S, causal=False, need_lse=save, use_tc=USE_TC_ATTENTION)
        y2 = ops.gemm(o2, wc["wo"], bias=wc["bo"])
        # --- MLP
        g, b_ = _ln(st, k + ".ln3", layer.final_layer_norm)
        x3, h3, mu3, rs3 = ops.add_layernorm(x2, y2, g, b_, rows=M, d=d, save_stats=save)
        w1, b1 = bf16_of(c, k + ".fc1.w", layer.fc1.weight), f32_of(c, k + ".fc1.b", layer.fc1.bias)
        w2, b2 = bf16_of(c, k + ".fc2.w", layer.fc2.weight), f32_of(c, k + ".fc2.b", layer.fc2.bias)
        if save:
            hpre = ops.gemm(h3, w1, bias=b1)                   # keep the pre-activation for gelu'
            a = ops.gelu_fwd(hpre)
        else:
            hpre, a = None, ops.gemm(h3, w1, bias=b1, act=1)
        y3 = ops.gemm(a, w2, bias=b2)
        if save:
            L.update(x1=x1, h1=h1, mu1=mu1, rs1=rs1, qkv=qkv, o1=o1, lse1=lse1, x2=x2, h2=h2, mu2=mu2, rs2=rs2, qc=qc,
                     kvc=kvc, o2=o2, lse2=lse2, x3=x3, h3=h3, mu3=mu3, rs3=rs3, hpre=hpre, a=a)
            ctx["layers"].append(L)
        x, y = x3, y3
    g, b_ = _ln(st, "ln_f", dec.layer_norm)
    xf, hf, muf, rsf = ops.add_layernorm(x, y, g, b_, rows=M, d=d, save_stats=save)
    if save:
        ctx.update(xf=xf, hf=hf, muf=muf, rsf=rsf)
    return hf, ctx


def lm_head(st: _State, hf):
    """proj_out (tied to embed_tokens, no bias; HF:models/whisper/modeling_whisper.py:1081) -> fp32 [M, ld>=V]."""
    dec = st.m
    V = dec.config.vocab_size
    Eb = bf16_of(st.cache, "E", dec.embed_tokens.weight)
    buf = torch.empty((hf.shape[0], ops.round_up(V, 8)), dtype=F32, device=hf.device)
    ops.gemm(hf, Eb, out=buf[:, :V])
    return buf


def _linear_bwd(dy, x_in, w_bf16, weight_p, bias_p, need_dx=True):
    """y = x W^T + b.  dy bf16 [M, N] (may be a column slice), x_in bf16 [M, K], w bf16 [N, K].
    Accumulates dW / db into the parameters' fp32 .grad, returns dx bf16 [M, K]."""
    if weight_p is not None and weight_p.requires_grad:
        ops.gemm(dy, x_in, a_mn=True, b_mn=True, out=_grad_buf(weight_p).reshape(weight_p.shape[0], -1), accumulate=True)
    if bias_p is not None and bias_p.requires_grad:
        ops.colsum(dy, out=_grad_buf(bias_p), accumulate=True)
    if need_dx:
        return ops.gemm(dy, w_bf16, b_mn=True)
    return None


def _ln_bwd(dy, x, mu, rs, ln, gamma_f32, dres, rows, d):
    lw, lb = ln.weight, ln.bias
    dg = _grad_buf(lw) if lw.requires_grad else None
    db = _grad_buf(lb) if lb.requires_grad else None
    return ops.layernorm_bwd(dy, x, mu, rs, gamma_f32, dres, dg, db, rows=rows, d=d)


def decoder_backward(st: _State, ctx, dlogits, want_denc=False):
    """dlogits: bf16 [M, ld>=V] gradient wrt the fp32 logits.  Accumulates every decoder parameter gradient into
    .grad (fp32) and returns d(encoder states) as fp32 [B*S, d] when want_denc."""
    dec = st.m
    cfg = dec.config
    d, H, V = cfg.d_model, cfg.decoder_attention_heads, cfg.vocab_size
    B, T, S = ctx["B"], ctx["T"], ctx["S"]
    M = B * T
    c = st.cache
    enc = ctx["enc"]
    dl = dlogits[:, :V]
    E = dec.embed_tokens.weight
    Eb = bf16_of(c, "E", E)
    # LM head: dhf = dlogits . E ; dE += dlogits^T . hf
    dhf = ops.gemm(dl, Eb, b_mn=True)
    if E.requires_grad:
        ops.gemm(dl, ctx["hf"], a_mn=True, b_mn=True, out=_grad_buf(E), accumulate=True)
    g, _ = _ln(st, "ln_f", dec.layer_norm)
    dx, dxb = _ln_bwd(dhf, ctx["xf"], ctx["muf"], ctx["rsf"], dec.layer_norm, g, None, M, d)
    denc = torch.zeros((B * S, d), dtype=F32, device=dl.device) if want_denc else None
    for i in reversed(range(len(dec.layers))):
        layer, L, k = dec.layers[i], ctx["layers"][i], f"l{i}"
        # ---- MLP: x_out = x3 + fc2(gelu(fc1(LN3(x3))))
        da = _linear_bwd(dxb, L["a"], bf16_of(c, k + ".fc2.w", layer.fc2.weight), layer.fc2.weight, layer.fc2.bias)
        dhpre = ops.gelu_bwd(da, L["hpre"])
        dh3 = _linear_bwd(dhpre, L["h3"], bf16_of(c, k + ".fc1.w", layer.fc1.weight), layer.fc1.weight, layer.fc1.bias)
        g, _ = _ln(st, k + ".ln3", layer.final_layer_norm)
        dx, dxb = _ln_bwd(dh3, L["x3"], L["mu3"], L["rs3"], layer.final_layer_norm, g, dx, M, d)
        # ---- cross-attention
        ca = layer.encoder_attn
        wc = _attn_weights(st, k + ".ca", ca, fuse_qkv=False)
        do2 = _linear_bwd(dxb, L["o2"], wc["wo"], ca.out_proj.weight, ca.out_proj.bias)
        dqc = torch.empty((M, d), dtype=BF16, device=dl.device)
        dkvc = torch.empty((B * S, 2 * d), dtype=BF16, device=dl.device)
        kvc = L["kvc"]
        ops.attention_bwd(L["qc"], kvc[:, :d], kvc[:, d:], L["o2"], do2, L["lse2"], B, H, T, S, False, dqc, dkvc[:, :d], dkvc[:, d:])
        dh2 = _linear_bwd(dqc, L["h2"], wc["wq"], ca.q_proj.weight, ca.q_proj.bias)
        _linear_bwd(dkvc[:, :d], enc, None, ca.k_proj.weight, None, need_dx=False)
        _linear_bwd(dkvc[:, d:], enc, None, ca.v_proj.weight, ca.v_proj.bias, need_dx=False)
        if want_denc:
            ops.gemm(dkvc, wc["wkv"], b_mn=True, out=denc, accumulate=True)
        g, _ = _ln(st, k + ".ln2", layer.encoder_attn_layer_norm)
        dx, dxb = _ln_bwd(dh2, L["x2"], L["mu2"], L["rs2"], layer.encoder_attn_layer_norm, g, dx, M, d)
        # ---- causal self-attention
        sa = layer.self_attn
        ws = _attn_weights(st, k + ".sa", sa, fuse_qkv=True)
        do1 = _linear_bwd(dxb, L["o1"], ws["wo"], sa.out_proj.weight, sa.out_proj.bias)
        qkv = L["qkv"]
        dqkv = torch.empty((M, 3 * d), dtype=BF16, device=dl.device)
        ops.attention_bwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], L["o1"], do1, L["lse1"], B, H, T, T, True,
                          dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:])
        _linear_bwd(dqkv[:, :d], L["h1"], None, sa.q_proj.weight, sa.q_proj.bias, need_dx=False)
        _linear_bwd(dqkv[:, d:2 * d], L["h1"], None, sa.k_proj.weight, None, need_dx=False)
        _linear_bwd(dqkv[:, 2 * d:], L["h1"], None, sa.v_proj.weight, sa.v_proj.bias, need_dx=False)
        dh1 = ops.gemm(dqkv, ws["wqkv"], b_mn=True)
        g, _ = _ln(st, k + ".ln1", layer.self_attn_layer_norm)
        dx, dxb = _ln_bwd(dh1, L["x1"], L["mu1"], L["rs1"], layer.self_attn_layer_norm, g, dx, M, d)
    # embeddings (dx is now the gradient wrt the embedding output)
    P = dec.embed_positions.weight
    dE = _grad_buf(E) if E.requires_grad else None
    dP = _grad_buf(P) if P.requires_grad else None
    if dE is not None or dP is not None:
        if dP is not None and dP.shape[0] < T:
            raise ValueError("embed_positions shorter than the sequence")
        ops.embed_bwd(ctx["ids"], dx, dE, dP, B, T, d, V, dec.embed_tokens.padding_idx if dec.embed_tokens.padding_idx is not None else -1)
    return denc


# =================================================================================================================
# model-level helpers and autograd bridges
def _check_trainable_dtypes(model):
    for n, p in model.named_parameters():
        if p.requires_grad and p.dtype != F32:
            raise TypeError(f"trainable parameter {n} is {p.dtype}; the student keeps fp32 master weights "
                            "(ref:training/run_distillation.py:995-1004 loads it in fp32)")


def encoder_is_trainable(model) -> bool:
    return any(p.requires_grad for n, p in model.model.encoder.named_parameters() if "embed_positions" not in n)


def run_encoder(model, input_features, enc_in, save=False):
    """Returns (bf16 [B*S, d] encoder states, S, ctx) from features or from caller-provided hidden states.
    save=True (trainable encoder, variant A) keeps the activations encoder_backward needs in ctx."""
    cfg = model.config
    if enc_in is not None:
        e = enc_in.reshape(-1, cfg.d_model)
        if e.dtype == F32:
            e = ops.cast_f32_to_bf16(e.contiguous())
        elif e.dtype != BF16:
            raise TypeError(f"encoder_outputs dtype {e.dtype} unsupported")
        return e.contiguous(), e.shape[0] // enc_in.shape[0], None
    if input_features is None:
        raise ValueError("input_features or encoder_outputs are required")
    with nvtx_range("dwb.encoder_forward"):
        out, ectx = encoder_forward(state_of(model.model.encoder), input_features, save=save)
    return out, cfg.max_source_positions, ectx


def backward_through_model(model, dctx, ectx, dlogits_bf16):
    """Decoder (+ encoder when ectx is given) backward from the bf16 logits gradient; gradients land in .grad."""
    with nvtx_range("dwb.decoder_backward"):
        denc = decoder_backward(state_of(model.model.decoder), dctx, dlogits_bf16, want_denc=ectx is not None)
    if ectx is not None:
        with nvtx_range("dwb.encoder_backward"):
            encoder_backward(state_of(model.model.encoder), ectx, ops.cast_f32_to_bf16(denc))


class ModelForwardFn(torch.autograd.Function):
    """logits = model(input_features | encoder states, decoder_input_ids).  The backward runs the decoder schedule and
    writes parameter gradients straight into .grad (fp32, accumulated), like autograd's AccumulateGrad would."""

    @staticmethod
    def run(model, input_features, decoder_input_ids, enc_in):
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in model.parameters())
        anchor = next((p for p in model.parameters() if p.requires_grad), None) if need_grad else None
        if need_grad:
            _check_trainable_dtypes(model)
            return ModelForwardFn.apply(model, input_features, decoder_input_ids, enc_in, anchor)
        with torch.no_grad():
            return ModelForwardFn.forward(None, model, input_features, decoder_input_ids, enc_in, None)

    @staticmethod
    def forward(ctx, model, input_features, decoder_input_ids, enc_in, anchor):
        cfg = model.config
        save = ctx is not None
        enc, S, ectx = run_encoder(model, input_features, enc_in, save=save and enc_in is None and encoder_is_trainable(model))
        B, T = decoder_input_ids.shape
        dst = state_of(model.model.decoder)
        hf, dctx = decoder_forward(dst, decoder_input_ids, enc, B, S, save)
        buf = lm_head(dst, hf)
        logits = buf.view(B, T, -1)[:, :, :cfg.vocab_size]
        if save:
            ctx.model, ctx.dctx, ctx.ectx = model, dctx, ectx
            ctx.mark_non_differentiable(enc)
        return logits, enc

    @staticmethod
    def backward(ctx, dlogits, _denc):
        model, cfg = ctx.model, ctx.model.config
        B, T, V = dlogits.shape
        ld = ops.round_up(V, 8)
        dl = torch.zeros((B * T, ld), dtype=BF16, device=dlogits.device)
        src = dlogits.reshape(B * T, V)
        if src.dtype != F32:
            src = src.float()
        # the cast kernel wants 16 B aligned fp32 rows: stage through a padded fp32 buffer when V % 4 != 0
        if V % 4 or src.stride(0) % 4 or src.data_ptr() % 16:
            pad = torch.zeros((B * T, ops.round_up(V, 4)), dtype=F32, device=dlogits.device)
            pad[:, :V].copy_(src)
            ops.cast_f32_to_bf16(pad, dl[:, :pad.shape[1]])
        else:
            ops.cast_f32_to_bf16(src, dl[:, :V])
        backward_through_model(model, ctx.dctx, ctx.ectx, dl)
        ctx.dctx = ctx.ectx = None
        return None, None, None, None, None


class CrossEntropyFn(torch.autograd.Function):
    """HF:models/whisper/modeling_whisper.py:1085-1088: CrossEntropyLoss(ignore_index=-100, mean) on fp32 logits."""

    @staticmethod
    def forward(ctx, logits, labels, vocab):
        B, T, V = logits.shape
        flat = logits.reshape(B * T, V)
        if flat.stride(0) % 4 or flat.data_ptr() % 16:
            buf = torch.empty((B * T, ops.round_up(V, 8)), dtype=F32, device=logits.device)
            buf[:, :V].copy_(flat)
            flat = buf[:, :V]
        full = torch.as_strided(flat, (B * T, flat.stride(0)), (flat.stride(0), 1))
        metrics, dl = ops.kd_loss(full, None, labels, V, 1.0, 1.0, 0.0, want_grad=ctx.needs_input_grad[0] if ctx else False)
        ctx.dl, ctx.shape = dl, (B, T, V)
        return metrics[0].clone()

    @staticmethod
    def backward(ctx, g):
        B, T, V = ctx.shape
        d = ops.cast_bf16_to_f32(ctx.dl).view(B, T, -1)[:, :, :V]
        return d * g, None, None
