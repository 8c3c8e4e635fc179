// HBM-bound kernels around the GEMMs: residual-add + LayerNorm (fwd/bwd), casts, conv im2col, embedding
// (fwd/bwd), bias gradients (column sums), GELU backward.  All are one pass over their operands with 16 B
// vector accesses; rows are d_model wide (<= 2048) so a warp owns a row and keeps it in registers.
#include "common.cuh"

namespace dwb {

constexpr int LN_MAX_VEC = 16;   // float4 per lane: d <= 16*4*32 = 2048

// ------------------------------------------------------------------------------------------------
// x_new = x_in[row % x_rows_mod] (+ y);  ln = LayerNorm(x_new) * gamma + beta
// HF:models/whisper/modeling_whisper.py:393-409 / :470-503 residual adds followed by the next pre-LN, and
// :623-625 (the conv stem output plus the positional table feeding layer 0's LN) are all this one pattern.
template <int NV, int MINB>
__global__ void __launch_bounds__(256, MINB) add_layernorm_kernel(const float* __restrict__ x_in, int x_rows_mod,
                                                            const bf16* __restrict__ y, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ x_out,
                                                            bf16* __restrict__ ln_out, float* __restrict__ mean_out,
                                                            float* __restrict__ rstd_out, int rows, int d, float eps, int reverse) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  // Rows are visited from the last to the first: the GEMM that produced x (TMA reduce-add epilogue) walks its tiles with m ascending,
  // so the END of x is what is still in the 126 MB L2 when this kernel starts, and the GEMM that consumes the normalised rows starts at
  // row 0 -- the rows this kernel writes last.
  const int row = reverse ? rows - 1 - warp : warp;
  const int src_row = x_rows_mod > 0 ? row % x_rows_mod : row;
  const float4* xr = reinterpret_cast<const float4*>(x_in + (int64_t)src_row * d);
  const uint2* yr = y ? reinterpret_cast<const uint2*>(y + (int64_t)row * d) : nullptr;
  const int nvec = d >> 2;
  float4 v[NV];   // NV = float4 per lane, sized to the row (register budget decides how many rows an SM keeps in flight)
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = i * 32 + lane;
    if (idx < nvec) {
      float4 a = xr[idx];
      if (yr) {
        const uint2 u = yr[idx];
        const float2 lo = unpack_bf16x2(u.x), hi = unpack_bf16x2(u.y);
        a.x += lo.x; a.y += lo.y; a.z += hi.x; a.w += hi.y;
      }
      v[i] = a;
      sum += a.x + a.y + a.z + a.w;
    }
  }
  const float mean = warp_sum(sum) / (float)d;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = i * 32 + lane;
    if (idx < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      sq += a * a + b * b + c * c + e * e;
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)d + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  float4* xo = x_out ? reinterpret_cast<float4*>(x_out + (int64_t)row * d) : nullptr;
  uint2* lo_ = ln_out ? reinterpret_cast<uint2*>(ln_out + (int64_t)row * d) : nullptr;
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = i * 32 + lane;
    if (idx < nvec) {
      if (xo) xo[idx] = v[i];
      if (lo_) {
        const float4 g = g4[idx], b = b4[idx];
        const float o0 = (v[i].x - mean) * rstd * g.x + b.x, o1 = (v[i].y - mean) * rstd * g.y + b.y;
        const float o2 = (v[i].z - mean) * rstd * g.z + b.z, o3 = (v[i].w - mean) * rstd * g.w + b.w;
        lo_[idx] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
      }
    }
  }
}

// LayerNorm backward.  dy: gradient wrt the LN output (bf16, from the dgrad GEMM); x: the fp32 LN input;
// dres: gradient arriving through the skip connection (fp32, nullable).  Writes dx = dres + LN'(dy) as fp32 and,
// optionally, as bf16 (the next dgrad/wgrad GEMM's operand); accumulates dgamma/dbeta with one atomic per
// column per CTA.
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const bf16* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ dres,
                                                            float* __restrict__ dx, bf16* __restrict__ dx_bf16,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta, int rows,
                                                            int d, int rows_per_cta) {
  extern __shared__ float red[];   // [8 warps][2][d]  -> reduced by warp 0.. at the end
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = d >> 2;
  float4 dg[LN_MAX_VEC], db[LN_MAX_VEC];
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC; ++i) dg[i] = db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int row_begin = blockIdx.x * rows_per_cta;
  const int row_end = min(rows, row_begin + rows_per_cta);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  for (int row = row_begin + warp; row < row_end; row += 8) {
    const float mu = mean[row], rs = rstd[row];
    const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)row * d);
    const uint2* dyr = reinterpret_cast<const uint2*>(dy + (int64_t)row * d);
    float4 xh[LN_MAX_VEC], gy[LN_MAX_VEC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_VEC; ++i) {
      const int idx = i * 32 + lane;
      if (idx < nvec) {
        const float4 a = xr[idx];
        const uint2 u = dyr[idx];
        const float2 lo = unpack_bf16x2(u.x), hi = unpack_bf16x2(u.y);
        const float4 g = g4[idx];
        const float4 h = make_float4((a.x - mu) * rs, (a.y - mu) * rs, (a.z - mu) * rs, (a.w - mu) * rs);
        const float4 t = make_float4(lo.x * g.x, lo.y * g.y, hi.x * g.z, hi.y * g.w);
        xh[i] = h; gy[i] = t;
        s1 += t.x + t.y + t.z + t.w;
        s2 += t.x * h.x + t.y * h.y + t.z * h.z + t.w * h.w;
        dg[i].x += lo.x * h.x; dg[i].y += lo.y * h.y; dg[i].z += hi.x * h.z; dg[i].w += hi.y * h.w;
        db[i].x += lo.x; db[i].y += lo.y; db[i].z += hi.x; db[i].w += hi.y;
      }
    }
    s1 = warp_sum(s1) / (float)d;
    s2 = warp_sum(s2) / (float)d;
    const float4* rr = dres ? reinterpret_cast<const float4*>(dres + (int64_t)row * d) : nullptr;
    float4* dxr = reinterpret_cast<float4*>(dx + (int64_t)row * d);
    uint2* dxb = dx_bf16 ? reinterpret_cast<uint2*>(dx_bf16 + (int64_t)row * d) : nullptr;
#pragma unroll
    for (int i = 0; i < LN_MAX_VEC; ++i) {
      const int idx = i * 32 + lane;
      if (idx < nvec) {
        float4 o;
        o.x = rs * (gy[i].x - s1 - xh[i].x * s2);
        o.y = rs * (gy[i].y - s1 - xh[i].y * s2);
        o.z = rs * (gy[i].z - s1 - xh[i].z * s2);
        o.w = rs * (gy[i].w - s1 - xh[i].w * s2);
        if (rr) { const float4 r = rr[idx]; o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
        dxr[idx] = o;
        if (dxb) dxb[idx] = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
      }
    }
  }
  // cross-warp reduction of dgamma / dbeta partials
  float4* red4 = reinterpret_cast<float4*>(red);
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC; ++i) {
    const int idx = i * 32 + lane;
    if (idx < nvec) {
      red4[(warp * 2 + 0) * nvec + idx] = dg[i];
      red4[(warp * 2 + 1) * nvec + idx] = db[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { a += red[(w * 2 + 0) * d + c]; b += red[(w * 2 + 1) * d + c]; }
    if (dgamma) atomicAdd(dgamma + c, a);
    if (dbeta) atomicAdd(dbeta + c, b);
  }
}

// ------------------------------------------------------------------------------------------------
// 2-D strided cast fp32 -> bf16 (rows x cols, cols % 4 == 0), optionally scaled
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, int64_t lds, bf16* __restrict__ dst, int64_t ldd, int rows,
                                     int cols, float scale) {
  const int64_t nvec = (int64_t)rows * (cols >> 2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / (cols >> 2)), c = (int)(i % (cols >> 2)) << 2;
    const float4 a = *reinterpret_cast<const float4*>(src + (int64_t)r * lds + c);
    *reinterpret_cast<uint2*>(dst + (int64_t)r * ldd + c) =
        make_uint2(pack_bf16x2(a.x * scale, a.y * scale), pack_bf16x2(a.z * scale, a.w * scale));
  }
}
__global__ void cast_bf16_f32_kernel(const bf16* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd, int rows,
                                     int cols) {
  const int64_t nvec = (int64_t)rows * (cols >> 2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / (cols >> 2)), c = (int)(i % (cols >> 2)) << 2;
    const uint2 u = *reinterpret_cast<const uint2*>(src + (int64_t)r * lds + c);
    const float2 lo = unpack_bf16x2(u.x), hi = unpack_bf16x2(u.y);
    *reinterpret_cast<float4*>(dst + (int64_t)r * ldd + c) = make_float4(lo.x, lo.y, hi.x, hi.y);
  }
}
// conv weight [O, C, 3] fp32 -> bf16 [O, 3*C] with column index k*C + c (matches the channels-last im2col rows)
__global__ void conv_weight_kc_kernel(const float* __restrict__ w, bf16* __restrict__ out, int O, int C) {
  const int64_t n = (int64_t)O * C * 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % 3), c = (int)((i / 3) % C), o = (int)(i / (3 * (int64_t)C));
    out[(int64_t)o * 3 * C + (int64_t)k * C + c] = __float2bfloat16_rn(w[i]);
  }
}
// inverse mapping for the weight gradient: g[O, 3*C] (k*C + c, fp32) -> dw[O, C, 3] (accumulate)
__global__ void conv_wgrad_kc_to_ck_kernel(const float* __restrict__ g, float* __restrict__ dw, int O, int C, int accumulate) {
  const int64_t n = (int64_t)O * C * 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % 3), c = (int)((i / 3) % C), o = (int)(i / (3 * (int64_t)C));
    const float v = g[(int64_t)o * 3 * C + (int64_t)k * C + c];
    dw[i] = accumulate ? dw[i] + v : v;
  }
}

// ------------------------------------------------------------------------------------------------
// conv1 im2col: mel [B, C, L] fp32 (time contiguous) -> rows [B*L, ld] bf16, column c*3 + k = mel[b, c, t + k - 1]
// (zero padded), columns >= 3*C zero.  HF:models/whisper/modeling_whisper.py:619 (conv1 k3 s1 p1).
__global__ void im2col_conv1_kernel(const float* __restrict__ mel, bf16* __restrict__ out, int B, int C, int L, int ld) {
  __shared__ float tile[32][34];     // [c][t + halo]
  const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
  for (int cc = ty; cc < 32; cc += 8) {
    const int c = c0 + cc;
    for (int tt = tx; tt < 34; tt += 32) {
      const int t = t0 + tt - 1;
      tile[cc][tt] = (c < C && t >= 0 && t < L) ? mel[((int64_t)b * C + c) * L + t] : 0.f;
    }
  }
  __syncthreads();
  // each thread writes 3 taps of one (t, c): consecutive c -> consecutive 6 B
  for (int tt = ty; tt < 32; tt += 8) {
    const int t = t0 + tt, c = c0 + tx;
    if (t < L && c < C) {
      bf16* o = out + ((int64_t)b * L + t) * ld + c * 3;
      o[0] = __float2bfloat16_rn(tile[tx][tt]);
      o[1] = __float2bfloat16_rn(tile[tx][tt + 1]);
      o[2] = __float2bfloat16_rn(tile[tx][tt + 2]);
    }
  }
  if (blockIdx.y == 0) {           // zero the padding columns once
    for (int tt = ty; tt < 32; tt += 8) {
      const int t = t0 + tt;
      if (t < L) for (int c = 3 * C + tx; c < ld; c += 32) out[((int64_t)b * L + t) * ld + c] = __float2bfloat16_rn(0.f);
    }
  }
}
// conv2 im2col on channels-last input x [B, L, d] bf16: row (b, t) = x[b, 2t-1 .. 2t+1, :] (3*d contiguous, zero
// for position -1).  HF:models/whisper/modeling_whisper.py:620 (conv2 k3 s2 p1).
__global__ void im2col_conv2_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int B, int L, int d) {
  const int Lo = L / 2;
  const int vec_per_row = 3 * d / 8;
  const int64_t n = (int64_t)B * Lo * vec_per_row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int vcol = (int)(i % vec_per_row);
    const int64_t row = i / vec_per_row;
    const int t = (int)(row % Lo), b = (int)(row / Lo);
    const int e = vcol * 8;                 // element offset inside the 3*d window
    const int pos = 2 * t - 1 + e / d;      // source time step
    uint4 v = make_uint4(0, 0, 0, 0);
    if (pos >= 0 && pos < L) v = *reinterpret_cast<const uint4*>(x + ((int64_t)b * L + pos) * d + (e % d));
    *reinterpret_cast<uint4*>(out + row * 3 * d + e) = v;
  }
}

// ------------------------------------------------------------------------------------------------
// decoder input embedding: x[b,t,:] = E[ids[b,t]] + P[t]        HF:models/whisper/modeling_whisper.py:738,755
template <typename TE>
__global__ void embed_fwd_kernel(const int64_t* __restrict__ ids, const TE* __restrict__ E, const TE* __restrict__ P,
                                 float* __restrict__ x, int rows, int T, int d, int vocab) {
  const int row = blockIdx.x;
  if (row >= rows) return;
  const int64_t id = ids[row];
  const int t = row % T;
  // nn.Embedding raises on an out-of-range id (device assert); here the row is poisoned with NaN so that the loss and
  // every gradient of the step become NaN instead of silently decoding token 0 -- never an out-of-bounds read
  const bool ok = id >= 0 && id < vocab;
  for (int c = threadIdx.x; c < d; c += blockDim.x)
    x[(int64_t)row * d + c] = ok ? (float)E[id * d + c] + (float)P[(int64_t)t * d + c] : __int_as_float(0x7fc00000);
}
// dE[ids] += dx (skipping padding_idx), dP[t] += dx
__global__ void embed_bwd_kernel(const int64_t* __restrict__ ids, const float* __restrict__ dx, float* __restrict__ dE,
                                 float* __restrict__ dP, int rows, int T, int d, int vocab, int padding_idx) {
  const int row = blockIdx.x;
  if (row >= rows) return;
  const int64_t id = ids[row];
  const int t = row % T;
  const bool tok_ok = dE != nullptr && id >= 0 && id < vocab && id != padding_idx;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float g = dx[(int64_t)row * d + c];
    if (tok_ok) atomicAdd(dE + id * d + c, g);
    if (dP) atomicAdd(dP + (int64_t)t * d + c, g);
  }
}

// ------------------------------------------------------------------------------------------------
// bias gradient: out[c] (+)= sum_r m[r, c]   (m bf16 [rows, ld])
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const bf16* __restrict__ m, int64_t ld, float* __restrict__ out, int rows,
                                                          int cols, int rows_per_cta) {
  // block = 32 (column pairs) x 8 (row lanes); each thread owns two adjacent columns
  __shared__ float red[8][64];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 2;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  float a = 0.f, b = 0.f;
  if (c < cols) {
    for (int r = r0 + threadIdx.y; r < r1; r += 8) {
      const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(m + (int64_t)r * ld + c));
      a += v.x; b += v.y;
    }
  }
  red[threadIdx.y][threadIdx.x * 2] = a;
  red[threadIdx.y][threadIdx.x * 2 + 1] = b;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { sa += red[w][threadIdx.x * 2]; sb += red[w][threadIdx.x * 2 + 1]; }
    atomicAdd(out + c, sa);
    if (c + 1 < cols) atomicAdd(out + c + 1, sb);
  }
}

// dh = da * gelu'(h)    (bf16, n % 8 == 0)
__global__ void gelu_bwd_kernel(const bf16* __restrict__ da, const bf16* __restrict__ h, bf16* __restrict__ dh, int64_t n) {
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += (int64_t)gridDim.x * blockDim.x * 8) {
    const uint4 a = *reinterpret_cast<const uint4*>(da + i);
    const uint4 b = *reinterpret_cast<const uint4*>(h + i);
    const uint32_t au[4] = {a.x, a.y, a.z, a.w}, bu[4] = {b.x, b.y, b.z, b.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 g = unpack_bf16x2(au[j]), x = unpack_bf16x2(bu[j]);
      o[j] = pack_bf16x2(g.x * gelu_erf_grad(x.x), g.y * gelu_erf_grad(x.y));
    }
    *reinterpret_cast<uint4*>(dh + i) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
// y = gelu(h) elementwise (bf16) -- used when a layer must keep the pre-activation for its backward
__global__ void gelu_fwd_kernel(const bf16* __restrict__ h, bf16* __restrict__ y, int64_t n) {
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += (int64_t)gridDim.x * blockDim.x * 8) {
    const uint4 b = *reinterpret_cast<const uint4*>(h + i);
    const uint32_t bu[4] = {b.x, b.y, b.z, b.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = unpack_bf16x2(bu[j]);
      o[j] = pack_bf16x2(gelu_erf(x.x), gelu_erf(x.y));
    }
    *reinterpret_cast<uint4*>(y + i) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// x *= *scale for a contiguous bf16 buffer (n % 8 == 0); every thread leaves at once when *scale == 1 -- the upstream
// gradient `loss.backward()` supplies -- so the common case costs one empty launch, not a pass over HBM
__global__ void __launch_bounds__(256) scale_bf16_dev_kernel(bf16* __restrict__ x, int64_t n8, const float* __restrict__ scale) {
  const float s = *scale;
  if (s == 1.0f) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    uint4 v = reinterpret_cast<uint4*>(x)[i];
    uint32_t* u = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(u[j]);
      u[j] = pack_bf16x2(f.x * s, f.y * s);
    }
    reinterpret_cast<uint4*>(x)[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// Label side of DataCollatorSpeechSeq2SeqWithPadding (ref:training/run_distillation.py:460-476) on the device:
//   tokens [B, L+1] (padded token ids), lengths [B]  ->  decoder_input_ids = tokens[:, :-1],
//   labels = tokens[:, 1:] with -100 on padding (t + 1 >= length) and on the prompt: everything before, and including, the
//   first <|startoftranscript|> found at a label index > 0 (torch.argmax of the boolean row returns the first hit, 0 if none).
__global__ void __launch_bounds__(128) collate_labels_kernel(const int64_t* __restrict__ tokens, const int* __restrict__ lengths, int L1,
                                                             int64_t sot, int64_t* __restrict__ dec_in, int64_t* __restrict__ labels) {
  __shared__ int s_first;
  const int b = blockIdx.x, L = L1 - 1, len = lengths[b];
  const int64_t* row = tokens + (int64_t)b * L1;
  if (threadIdx.x == 0) s_first = 0x7fffffff;
  __syncthreads();
  int first = 0x7fffffff;
  for (int t = threadIdx.x; t < L; t += blockDim.x)
    if (t + 1 < len && row[t + 1] == sot) { first = t; break; }        // per-thread indices increase: the first hit is its smallest
  if (first != 0x7fffffff) atomicMin(&s_first, first);
  __syncthreads();
  int bos = s_first == 0x7fffffff ? 0 : s_first;
  if (bos > 0) bos += 1;
  for (int t = threadIdx.x; t < L; t += blockDim.x) {
    dec_in[(int64_t)b * L + t] = row[t];
    labels[(int64_t)b * L + t] = (t + 1 < len && t >= bos) ? row[t + 1] : -100;
  }
}

static inline int grid_for(int64_t work_items, int threads) {
  int64_t g = ceil_div64(work_items, threads);
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace dwb

using namespace dwb;

extern "C" int dwb_add_layernorm(const float* x_in, int x_rows_mod, const void* y_bf16, const float* gamma, const float* beta,
                                 float* x_out, void* ln_out_bf16, float* mean_out, float* rstd_out, int rows, int d, float eps,
                                 void* stream) {
  DWB_CHECK_ARG(x_in && gamma && beta, "dwb_add_layernorm: null operand");
  DWB_CHECK_ARG(rows > 0 && d > 0 && (d % 4) == 0 && d <= LN_MAX_VEC * 128, "dwb_add_layernorm: d=%d unsupported", d);
  const int nvec = ceil_div(d, 128);
  const dim3 grid(ceil_div(rows, 8));
  cudaStream_t st = (cudaStream_t)stream;
#define DWB_LN_LAUNCH(NV, MINB)                                                                                              \
  add_layernorm_kernel<NV, MINB><<<grid, 256, 0, st>>>(x_in, x_rows_mod, (const bf16*)y_bf16, gamma, beta, x_out, (bf16*)ln_out_bf16, \
                                                       mean_out, rstd_out, rows, d, eps, reverse)
  static const int wide_only = [] { const char* e = getenv("DWB_LN_WIDE"); return e ? atoi(e) : 0; }();   // A/B switch for the microbench
  const int reverse = dwb_row_walk_reverse();
  if (wide_only) DWB_LN_LAUNCH(16, 2);
  else if (nvec <= 3) DWB_LN_LAUNCH(3, 4);
  else if (nvec <= 4) DWB_LN_LAUNCH(4, 4);
  else if (nvec <= 6) DWB_LN_LAUNCH(6, 4);
  else if (nvec <= 8) DWB_LN_LAUNCH(8, 4);
  else if (nvec <= 10) DWB_LN_LAUNCH(10, 4);
  else DWB_LN_LAUNCH(16, 2);
#undef DWB_LN_LAUNCH
  DWB_LAUNCH_OK();
  return DWB_OK;
}

extern "C" int dwb_layernorm_bwd(const void* dy_bf16, const float* x, const float* mean, const float* rstd, const float* gamma,
                                 const float* dres, float* dx, void* dx_bf16, float* dgamma, float* dbeta, int rows, int d,
                                 void* stream) {
  DWB_CHECK_ARG(dy_bf16 && x && mean && rstd && gamma && dx, "dwb_layernorm_bwd: null operand");
  DWB_CHECK_ARG(rows > 0 && (d % 4) == 0 && d <= LN_MAX_VEC * 128, "dwb_layernorm_bwd: d=%d unsupported", d);
  int ctas = kNumSMs * 2;
  int rows_per_cta = ceil_div(rows, ctas);
  if (rows_per_cta < 8) rows_per_cta = 8;
  ctas = ceil_div(rows, rows_per_cta);
  const int smem = 8 * 2 * d * (int)sizeof(float);
  static int smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    DWB_CUDA_OK(cudaFuncSetAttribute(layernorm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    smem_set = smem;
  }
  layernorm_bwd_kernel<<<ctas, 256, smem, (cudaStream_t)stream>>>((const bf16*)dy_bf16, x, mean, rstd, gamma, dres, dx, (bf16*)dx_bf16,
                                                                  dgamma, dbeta, rows, d, rows_per_cta);
  DWB_LAUNCH_OK();
  return DWB_OK;
}

extern "C" int dwb_cast_f32_to_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int rows, int cols, float scale,
                                    void* stream) {
  DWB_CHECK_ARG(src && dst && rows > 0 && cols > 0 && (cols % 4) == 0 && (lds % 4) == 0 && (ldd % 4) == 0,
                "dwb_cast_f32_to_bf16: bad shape/alignment rows=%d cols=%d", rows, cols);
  cast_f32_bf16_kernel<<<grid_for((int64_t)rows * cols / 4, 256), 256, 0, (cudaStream_t)stream>>>(src, lds, (bf16*)dst, ldd, rows, cols,
                                                                                                  scale);
  DWB_LAUNCH_OK();
  return DWB_OK;
}
extern "C" int dwb_collate_labels(const int64_t* tokens, const int* lengths, int B, int L1, int64_t decoder_start_token_id,
                                  int64_t* decoder_input_ids, int64_t* labels, void* stream) {
  DWB_CHECK_ARG(tokens && lengths && decoder_input_ids && labels && B > 0 && L1 >= 2, "dwb_collate_labels: bad args");
  collate_labels_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(tokens, lengths, L1, decoder_start_token_id, decoder_input_ids, labels);
  DWB_LAUNCH_OK();
  return DWB_OK;
}
extern "C" int dwb_scale_bf16_dev(void* x_bf16, int64_t n, const float* scale_dev, void* stream) {
  DWB_CHECK_ARG(x_bf16 && scale_dev && n > 0 && (n % 8) == 0 && (reinterpret_cast<uintptr_t>(x_bf16) & 15) == 0,
                "dwb_scale_bf16_dev: bad args (n=%lld must be a multiple of 8, 16 B aligned)", (long long)n);
  scale_bf16_dev_kernel<<<grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((bf16*)x_bf16, n / 8, scale_dev);
  DWB_LAUNCH_OK();
  return DWB_OK;
}
extern "C" int dwb_cast_bf16_to_f32(const void* src, int64_t lds, float* dst, int64_t ldd, int rows, int cols, void* stream) {
  DWB_CHECK_ARG(src && dst && rows > 0 && cols > 0 && (cols % 4) == 0 && (lds % 4) == 0 && (ldd % 4) == 0,
                "dwb_cast_bf16_to_f32: bad shape/alignment");
  cast_bf16_f32_kernel<<<grid_for((int64_t)rows * cols / 4, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)src, lds, dst, ldd, rows,
                                                                                                  cols);
  DWB_LAUNCH_OK();
  return DWB_OK;
}
extern "C" int dwb_conv_weight_to_kc_bf16(const float* w, void* out_bf16, int O, int C, void* stream) {
  DWB_CHECK_ARG(w && out_bf16 && O > 0 && C > 0, "dwb_conv_weight_to_kc_bf16: bad args");
  conv_weight_kc_kernel<<<grid_for((int64_t)O * C * 3, 256), 256, 0, (cudaStream_t)stream>>>(w, (bf16*)out_bf16, O, C);
  DWB_LAUNCH_OK();
  return DWB_OK;
}
extern "C" int dwb_conv_wgrad_kc_to_ck(const float* g, float* dw, int O, int C, int accumulate, void* stream) {
  DWB_CHECK_ARG(g && dw && O > 0 && C > 0, "dwb_conv_wgrad_kc_to_ck: bad args");
  conv_wgrad_kc_to_ck_kernel<<<grid_for((int64_t)O * C * 3, 256), 256, 0, (cudaStream_t)stream>>>(g, dw, O, C, accumulate);
  DWB_LAUNCH_OK();
  return DWB_OK;
}
extern "C" int dwb_im2col_conv1(const float* mel, void* out_bf16, int B, int C, int L, int ld, void* stream) {
  DWB_CHECK_ARG(mel && out_bf16 && B > 0 && C > 0 && L > 0 && ld >= 3 * C && (ld % 8) == 0, "dwb_im2col_conv1: bad args");
  dim3 grid(ceil_div(L, 32), ceil_div(C, 32), B), block(32, 8);
  im2col_conv1_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(mel, (bf16*)out_bf16, B, C, L, ld);
  DWB_LAUNCH_OK();
  return DWB_OK;
}
extern "C" int dwb_im2col_conv2(const void* x_bf16, void* out_bf16, int B, int L, int d, void* stream) {
  DWB_CHECK_ARG(x_bf16 && out_bf16 && B > 0 && L > 0 && (L % 2) == 0 && (d % 8) == 0, "dwb_im2col_conv2: bad args");
  im2col_conv2_kernel<<<grid_for((int64_t)B * (L / 2) * (3 * d / 8), 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x_bf16,
                                                                                                           (bf16*)out_bf16, B, L, d);
  DWB_LAUNCH_OK();
  return DWB_OK;
}
extern "C" int dwb_embed_fwd(const int64_t* ids, const void* E, const void* P, int table_is_f32, float* x, int B, int T, int d,
                             int vocab, void* stream) {
  DWB_CHECK_ARG(ids && E && P && x && B > 0 && T > 0 && d > 0, "dwb_embed_fwd: bad args");
  if (table_is_f32)
    embed_fwd_kernel<float><<<B * T, 256, 0, (cudaStream_t)stream>>>(ids, (const float*)E, (const float*)P, x, B * T, T, d, vocab);
  else
    embed_fwd_kernel<bf16><<<B * T, 256, 0, (cudaStream_t)stream>>>(ids, (const bf16*)E, (const bf16*)P, x, B * T, T, d, vocab);
  DWB_LAUNCH_OK();
  return DWB_OK;
}
extern "C" int dwb_embed_bwd(const int64_t* ids, const float* dx, float* dE, float* dP, int B, int T, int d, int vocab,
                             int padding_idx, void* stream) {
  DWB_CHECK_ARG(ids && dx && B > 0 && T > 0 && d > 0, "dwb_embed_bwd: bad args");
  embed_bwd_kernel<<<B * T, 256, 0, (cudaStream_t)stream>>>(ids, dx, dE, dP, B * T, T, d, vocab, padding_idx);
  DWB_LAUNCH_OK();
  return DWB_OK;
}
extern "C" int dwb_colsum_bf16(const void* m_bf16, int64_t ld, float* out, int rows, int cols, int accumulate, void* stream) {
  DWB_CHECK_ARG(m_bf16 && out && rows > 0 && cols > 0 && (ld % 2) == 0, "dwb_colsum_bf16: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) DWB_CUDA_OK(cudaMemsetAsync(out, 0, (size_t)cols * sizeof(float), st));
  const int col_blocks = ceil_div(cols, 64);
  int row_blocks = ceil_div(kNumSMs * 4, col_blocks);
  int rows_per_cta = ceil_div(rows, row_blocks);
  if (rows_per_cta < 64) rows_per_cta = 64;
  row_blocks = ceil_div(rows, rows_per_cta);
  dim3 grid(col_blocks, row_blocks), block(32, 8);
  colsum_bf16_kernel<<<grid, block, 0, st>>>((const bf16*)m_bf16, ld, out, rows, cols, rows_per_cta);
  DWB_LAUNCH_OK();
  return DWB_OK;
}
extern "C" int dwb_gelu_bwd(const void* da, const void* h, void* dh, int64_t n, void* stream) {
  DWB_CHECK_ARG(da && h && dh && n > 0 && (n % 8) == 0, "dwb_gelu_bwd: bad args");
  gelu_bwd_kernel<<<grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)da, (const bf16*)h, (bf16*)dh, n);
  DWB_LAUNCH_OK();
  return DWB_OK;
}
extern "C" int dwb_gelu_fwd(const void* h, void* y, int64_t n, void* stream) {
  DWB_CHECK_ARG(h && y && n > 0 && (n % 8) == 0, "dwb_gelu_fwd: bad args");
  gelu_fwd_kernel<<<grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)h, (bf16*)y, n);
  DWB_LAUNCH_OK();
  return DWB_OK;
}

// ------------------------------------------------------------------------------------------------
// conv2 input gradient: col2im of the im2col-space gradient g [B*L/2, 3*d] (column k*d + c) back onto the channels-last
// conv1 activation grid [B, L, d], fused with conv1's GELU backward:  dpre1 = col2im(g) * gelu'(pre1).
// Position p receives tap k=1 of output t=p/2 (p even) or taps k=0 of t=(p+1)/2 and k=2 of t=(p-1)/2 (p odd).
namespace dwb {
__global__ void col2im_conv2_gelu_bwd_kernel(const bf16* __restrict__ g, const bf16* __restrict__ pre1, bf16* __restrict__ out, int B,
                                             int L, int d) {
  const int Lo = L / 2;
  const int vec = d / 8;
  const int64_t n = (int64_t)B * L * vec;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % vec) * 8;
    const int64_t row = i / vec;
    const int p = (int)(row % L), b = (int)(row / L);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    auto add = [&](int t, int k) {
      if (t < 0 || t >= Lo) return;
      const uint4 u = *reinterpret_cast<const uint4*>(g + ((int64_t)b * Lo + t) * 3 * d + (int64_t)k * d + c);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(w[j]);
        acc[2 * j] += f.x; acc[2 * j + 1] += f.y;
      }
    };
    if ((p & 1) == 0) {
      add(p >> 1, 1);
    } else {
      add((p + 1) >> 1, 0);
      add((p - 1) >> 1, 2);
    }
    const uint4 pu = *reinterpret_cast<const uint4*>(pre1 + row * d + c);
    const uint32_t pw[4] = {pu.x, pu.y, pu.z, pu.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = unpack_bf16x2(pw[j]);
      o[j] = pack_bf16x2(acc[2 * j] * gelu_erf_grad(x.x), acc[2 * j + 1] * gelu_erf_grad(x.y));
    }
    *reinterpret_cast<uint4*>(out + row * d + c) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
}  // namespace dwb

extern "C" int dwb_col2im_conv2_gelu_bwd(const void* g_bf16, const void* pre1_bf16, void* out_bf16, int B, int L, int d, void* stream) {
  DWB_CHECK_ARG(g_bf16 && pre1_bf16 && out_bf16 && B > 0 && L > 0 && (L % 2) == 0 && (d % 8) == 0, "dwb_col2im_conv2_gelu_bwd: bad args");
  dwb::col2im_conv2_gelu_bwd_kernel<<<dwb::grid_for((int64_t)B * L * (d / 8), 256), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)g_bf16, (const bf16*)pre1_bf16, (bf16*)out_bf16, B, L, d);
  DWB_LAUNCH_OK();
  return DWB_OK;
}
