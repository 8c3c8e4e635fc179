// Single-token decoder step for greedy generation (the eval loop's `model.generate`, ref:training/run_distillation.py:1524-1528,
// and the pseudo-labelling loop, ref:training/run_pseudo_labelling.py:861-927): what HF does with a KV cache
// (HF:models/whisper/modeling_whisper.py:315-340, EncoderDecoderCache) plus the greedy token pick of
// HF:generation/utils.py `_sample` (arg-max, finished rows padded, EOS bookkeeping) and the two Whisper logits processors
// (HF:generation/logits_process.py SuppressTokensLogitsProcessor / SuppressTokensAtBeginLogitsProcessor) as additive biases.
//
// Everything position-dependent is read from DEVICE memory (`pos_dev`), so one captured CUDA graph of the whole step is
// replayed for every token: embedding of seq[:, pos], per layer {LN, QKV GEMM, attention over the cache (this file), ...},
// LM head, pick -> seq[:, pos + 1], advance.
//
// All kernels here are HBM-bound by construction (one query row per (batch, head)): the attention kernel reads every cached
// K and V row exactly once (2 * len * 128 B per head), which is its algorithmic traffic.
#include "common.cuh"

namespace dwb {

// ------------------------------------------------------------------------------------------------
// x[b, :] = E[seq[b, pos]] + P[pos]
template <typename TE>
__global__ void embed_decode_kernel(const int64_t* __restrict__ seq, int seq_ld, const int* __restrict__ pos_dev, const TE* __restrict__ E,
                                    const TE* __restrict__ P, float* __restrict__ x, int d, int vocab) {
  const int b = blockIdx.x;
  const int pos = *pos_dev;
  const int64_t id = seq[(int64_t)b * seq_ld + pos];
  const bool ok = id >= 0 && id < vocab;
  for (int c = threadIdx.x; c < d; c += blockDim.x)
    x[(int64_t)b * d + c] = ok ? (float)E[id * d + c] + (float)P[(int64_t)pos * d + c] : __int_as_float(0x7fc00000);
}

// ------------------------------------------------------------------------------------------------
// One query row per (head, batch): o = softmax(scale * q K^T) V over cache rows [0, len).
//   self-attention (k_new != null): the step's own k / v row is first appended to the cache at row pos, len = pos + 1
//   cross-attention (k_new == null): len = fixed_len (the encoder positions), the cache is read-only
// Phase 1: thread t scores keys t, t+128, ... (a K row of one head is 128 B: eight 16 B loads, q in registers)
// Phase 2: block max / sum of exp2
// Phase 3: warp w accumulates keys w, w+4, ... ; lane l owns output dims 2l, 2l+1 (a V row of one head = one 128 B request)
constexpr int AD_THREADS = 128;      // (8 warps measured no better: 5.97 vs 4.8-5.5 ms per teacher token step)

__global__ void __launch_bounds__(AD_THREADS) attn_decode_kernel(const bf16* __restrict__ q, int64_t ldq, const bf16* __restrict__ k_new,
                                                                 const bf16* __restrict__ v_new, int64_t ld_new, bf16* __restrict__ k_cache,
                                                                 bf16* __restrict__ v_cache, int64_t ld_cache, int cache_rows,
                                                                 bf16* __restrict__ o, int64_t ldo, int fixed_len,
                                                                 const int* __restrict__ pos_dev, float scale_log2) {
  extern __shared__ float s_p[];                 // [len] scores -> probabilities
  __shared__ float s_red[AD_THREADS / 32];
  __shared__ float s_acc[AD_THREADS / 32][64];
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int len = fixed_len;
  bf16* kc = k_cache + (int64_t)b * cache_rows * ld_cache + h * 64;
  bf16* vc = v_cache + (int64_t)b * cache_rows * ld_cache + h * 64;
  if (k_new != nullptr) {
    const int pos = *pos_dev;
    len = pos + 1;
    if (tid < 16) {                              // append this step's key / value row (2 x 128 B) to the cache
      const bf16* src = (tid < 8 ? k_new : v_new) + (int64_t)b * ld_new + h * 64 + (tid & 7) * 8;
      bf16* dst = (tid < 8 ? kc : vc) + (int64_t)pos * ld_cache + (tid & 7) * 8;
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
    }
    __syncthreads();
  }
  // q * scale * log2(e) in registers
  float qf[64];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(q + (int64_t)b * ldq + h * 64);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 u = qp[i];
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(w[j]);
        qf[i * 8 + j * 2] = f.x * scale_log2;
        qf[i * 8 + j * 2 + 1] = f.y * scale_log2;
      }
    }
  }
  float mx = -INFINITY;
  for (int k = tid; k < len; k += AD_THREADS) {
    const uint4* kp = reinterpret_cast<const uint4*>(kc + (int64_t)k * ld_cache);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 u = kp[i];
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(w[j]);
        s = fmaf(qf[i * 8 + j * 2], f.x, s);
        s = fmaf(qf[i * 8 + j * 2 + 1], f.y, s);
      }
    }
    s_p[k] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = s_red[0];
#pragma unroll
  for (int w = 1; w < AD_THREADS / 32; ++w) mx = fmaxf(mx, s_red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int k = tid; k < len; k += AD_THREADS) {
    const float e = fast_exp2(s_p[k] - mx);
    s_p[k] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < AD_THREADS / 32; ++w) tot += s_red[w];
  const float inv = 1.f / tot;
  float a0 = 0.f, a1 = 0.f;
  int k = warp;
  constexpr int NW = AD_THREADS / 32, UNR = 16;
  for (; k + (UNR - 1) * NW < len; k += UNR * NW) {       // 16 independent 128 B requests in flight per warp: the loop is latency-bound
    uint32_t u[UNR];
    float pr[UNR];
#pragma unroll
    for (int i = 0; i < UNR; ++i) {
      const int kk = k + i * NW;
      u[i] = *reinterpret_cast<const uint32_t*>(vc + (int64_t)kk * ld_cache + 2 * lane);
      pr[i] = s_p[kk];
    }
#pragma unroll
    for (int i = 0; i < UNR; ++i) {
      const float2 f = unpack_bf16x2(u[i]);
      a0 = fmaf(pr[i], f.x, a0);
      a1 = fmaf(pr[i], f.y, a1);
    }
  }
  for (; k + 3 * NW < len; k += 4 * NW) {
    uint32_t u[4];
    float pr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kk = k + i * NW;
      u[i] = *reinterpret_cast<const uint32_t*>(vc + (int64_t)kk * ld_cache + 2 * lane);
      pr[i] = s_p[kk];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = unpack_bf16x2(u[i]);
      a0 = fmaf(pr[i], f.x, a0);
      a1 = fmaf(pr[i], f.y, a1);
    }
  }
  for (; k < len; k += NW) {
    const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(vc + (int64_t)k * ld_cache + 2 * lane));
    const float pr = s_p[k];
    a0 = fmaf(pr, f.x, a0);
    a1 = fmaf(pr, f.y, a1);
  }
  s_acc[warp][2 * lane] = a0;
  s_acc[warp][2 * lane + 1] = a1;
  __syncthreads();
  if (tid < 32) {
    float r0 = 0.f, r1 = 0.f;
#pragma unroll
    for (int w = 0; w < AD_THREADS / 32; ++w) { r0 += s_acc[w][2 * tid]; r1 += s_acc[w][2 * tid + 1]; }
    *reinterpret_cast<uint32_t*>(o + (int64_t)b * ldo + h * 64 + 2 * tid) = pack_bf16x2(r0 * inv, r1 * inv);
  }
}

// ------------------------------------------------------------------------------------------------
// Skinny GEMM for the decode step: C[M, N] = act(X[M, K] . W[N, K]^T + bias) with M = the batch (16 .. 64 rows).  The projection is
// weight-bandwidth bound (every weight is used by M rows only): the persistent 128-row tcgen05 tiles give such a problem 10 - 40 CTAs
// and 11 - 14 us; here N / 8 CTAs (160 - 640) of four warps stream the weights once from HBM, take the activations through L1 (all
// CTAs read the same X), and multiply with mma.sync.m16n8k16 (bf16, fp32 accumulate).
__device__ __forceinline__ void mma_m16n8k16_bf16(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Layout trick: the reduction index may be visited in any order as long as A and B agree, so lane (g, t) takes the 16 consecutive
// k of its row that start at 16 t inside each 64-wide chunk (two 16 B loads) and feeds them to four k-steps; every load is a full
// 32 B sector and a warp has all the weights of its k-range in flight at once.  A CTA = 8 output columns x 4 k-ranges (4 warps),
// partial sums combined through shared memory.
template <int MT>      // 16-row tiles of X: M = 16 * MT
__global__ void __launch_bounds__(128) gemm_skinny_kernel(const bf16* __restrict__ x, int64_t ldx, const bf16* __restrict__ w, int64_t ldw,
                                                         const float* __restrict__ bias, void* __restrict__ c, int64_t ldc, int c_f32,
                                                         int N, int K, int act) {
  __shared__ float s_part[3][MT][4][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * 8;
  float acc[MT][4];
#pragma unroll
  for (int m = 0; m < MT; ++m) { acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f; }
  const int kq = K >> 2;                                             // this warp's k-range: [warp * kq, (warp + 1) * kq), kq % 64 == 0
  const bf16* wrow = w + (int64_t)(n0 + g) * ldw + warp * kq + 16 * t;
  const bf16* xrow = x + (int64_t)g * ldx + warp * kq + 16 * t;
  constexpr int UNR = 5;                                             // chunks of 64 k with their weight loads in flight together
  for (int kc = 0; kc < kq; kc += 64 * UNR) {
    uint4 b[UNR][2];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (kc + 64 * u < kq) {
        b[u][0] = *reinterpret_cast<const uint4*>(wrow + kc + 64 * u);
        b[u][1] = *reinterpret_cast<const uint4*>(wrow + kc + 64 * u + 8);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (kc + 64 * u < kq) {
        const uint32_t bw[8] = {b[u][0].x, b[u][0].y, b[u][0].z, b[u][0].w, b[u][1].x, b[u][1].y, b[u][1].z, b[u][1].w};
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const bf16* xr = xrow + (int64_t)(16 * m) * ldx + kc + 64 * u;
          const uint4 a00 = *reinterpret_cast<const uint4*>(xr), a01 = *reinterpret_cast<const uint4*>(xr + 8);
          const uint4 a10 = *reinterpret_cast<const uint4*>(xr + 8 * ldx), a11 = *reinterpret_cast<const uint4*>(xr + 8 * ldx + 8);
          const uint32_t r0[8] = {a00.x, a00.y, a00.z, a00.w, a01.x, a01.y, a01.z, a01.w};      // row g:     this lane's 16 k
          const uint32_t r1[8] = {a10.x, a10.y, a10.z, a10.w, a11.x, a11.y, a11.z, a11.w};      // row g + 8
#pragma unroll
          for (int st = 0; st < 4; ++st) {                       // k-step st uses the lane's elements 4 st .. 4 st + 3
            const uint32_t af[4] = {r0[2 * st], r1[2 * st], r0[2 * st + 1], r1[2 * st + 1]};
            const uint32_t bf[2] = {bw[2 * st], bw[2 * st + 1]};
            mma_m16n8k16_bf16(acc[m], af, bf);
          }
        }
      }
    }
  }
  if (warp > 0) {
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int i = 0; i < 4; ++i) s_part[warp - 1][m][i][lane] = acc[m][i];
  }
  __syncthreads();
  if (warp > 0) return;
  // C fragment: rows g, g + 8; columns n0 + 2t, n0 + 2t + 1
  const int col = n0 + 2 * t;
  const float b0 = bias ? bias[col] : 0.f, b1 = bias ? bias[col + 1] : 0.f;
#pragma unroll
  for (int m = 0; m < MT; ++m) {
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[m][i] += (s_part[0][m][i][lane] + s_part[1][m][i][lane]) + s_part[2][m][i][lane];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float v0 = acc[m][2 * h] + b0, v1 = acc[m][2 * h + 1] + b1;
      if (act == 1) { v0 = gelu_erf(v0); v1 = gelu_erf(v1); }
      const int64_t row = 16 * m + g + 8 * h;
      if (c_f32) *reinterpret_cast<float2*>(reinterpret_cast<float*>(c) + row * ldc + col) = make_float2(v0, v1);
      else *reinterpret_cast<uint32_t*>(reinterpret_cast<bf16*>(c) + row * ldc + col) = pack_bf16x2(v0, v1);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Greedy pick for row b after the step at position pos produced the logits of position pos + 1:
//   pos + 1 < prompt_len : the next token is the prompt's (already in seq), nothing is written
//   else next = finished[b] ? pad : argmax_v(logits[b, v] + bias_all[v] + (pos + 1 == begin_pos ? bias_begin[v] : 0));
//        seq[b, pos + 1] = next; finished[b] |= next == eos          (lowest index wins ties, like torch.argmax)
constexpr int GP_THREADS = 256;
__global__ void __launch_bounds__(GP_THREADS) greedy_pick_kernel(const float* __restrict__ logits, int64_t ld, int vocab,
                                                                 const float* __restrict__ bias_all, const float* __restrict__ bias_begin,
                                                                 int begin_pos, int64_t* __restrict__ seq, int seq_ld, int prompt_len,
                                                                 int* __restrict__ finished, int64_t eos, int64_t pad,
                                                                 const int* __restrict__ pos_dev) {
  __shared__ float s_v[GP_THREADS / 32];
  __shared__ int s_i[GP_THREADS / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int nxt = *pos_dev + 1;
  if (nxt < prompt_len || nxt >= seq_ld) return;
  if (finished[b]) {
    if (tid == 0) seq[(int64_t)b * seq_ld + nxt] = pad;
    return;
  }
  const float* row = logits + (int64_t)b * ld;
  const bool at_begin = bias_begin != nullptr && nxt == begin_pos;
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  for (int c = tid; c < vocab; c += GP_THREADS) {
    float v = row[c];
    if (bias_all) v += bias_all[c];
    if (at_begin) v += bias_begin[c];
    if (v > best) { best = v; best_i = c; }       // strided scan: within a thread indices increase, so '>' keeps the lowest
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
  }
  if ((tid & 31) == 0) { s_v[tid >> 5] = best; s_i[tid >> 5] = best_i; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < GP_THREADS / 32; ++w)
      if (s_v[w] > best || (s_v[w] == best && s_i[w] < best_i)) { best = s_v[w]; best_i = s_i[w]; }
    if (best_i == 0x7fffffff) best_i = 0;          // all -inf / NaN row
    seq[(int64_t)b * seq_ld + nxt] = best_i;
    if ((int64_t)best_i == eos) finished[b] = 1;
  }
}

// ------------------------------------------------------------------------------------------------
// Greedy pick under Whisper's timestamp rules (HF:generation/logits_process.py WhisperTimeStampLogitsProcessor, applied after the two
// suppress processors like HF:models/whisper/generation_whisper.py:1774-1800 orders them), for `return_timestamps=True` -- the
// reference's recommended pseudo-labelling mode (ref:training/README.md:130,148).  With g = the tokens generated so far (after the
// initial tokens), ts = ids >= ts_begin (= <|notimestamps|> + 1):
//   <|notimestamps|> is never emitted; after a timestamp that closes a pair (or opens the transcript) no timestamp may follow, after a
//   text token followed by one timestamp no text below EOS may follow; timestamps never decrease (and <|0.00|> is not repeated);
//   the first generated token is a timestamp no later than max_initial_timestamp_index; and whenever the probability mass of all
//   timestamps exceeds the most probable text token, a timestamp is forced.  All masks are index ranges, so one scan of the row that
//   tracks (best text token, best timestamp, max text logit, log-sum-exp of the timestamp logits) decides the pick.
struct TimestampRules {
  int ts_begin, no_ts, eos, max_initial;     // max_initial < 0: no limit
};
__global__ void __launch_bounds__(GP_THREADS) greedy_pick_ts_kernel(const float* __restrict__ logits, int64_t ld, int vocab,
                                                                    const float* __restrict__ bias_all, const float* __restrict__ bias_begin,
                                                                    int begin_pos, int64_t* __restrict__ seq, int seq_ld, int prompt_len,
                                                                    int* __restrict__ finished, int64_t eos, int64_t pad,
                                                                    const int* __restrict__ pos_dev, TimestampRules r) {
  __shared__ float s_f[4][GP_THREADS / 32];
  __shared__ int s_i[3][GP_THREADS / 32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nxt = *pos_dev + 1;
  if (nxt < prompt_len || nxt >= seq_ld) return;
  if (finished[b]) {
    if (tid == 0) seq[(int64_t)b * seq_ld + nxt] = pad;
    return;
  }
  const int64_t* srow = seq + (int64_t)b * seq_ld;
  // ---- state of the generated part g = srow[begin_pos .. nxt)
  const int n_gen = nxt - begin_pos;
  const bool last_ts = n_gen >= 1 && srow[nxt - 1] >= r.ts_begin;
  const bool penult_ts = n_gen < 2 || srow[nxt - 2] >= r.ts_begin;
  int last_idx = -1;                                        // index of the last timestamp token in g
  for (int i = begin_pos + tid; i < nxt; i += GP_THREADS)
    if (srow[i] >= r.ts_begin) last_idx = i;                // per-thread indices increase
  last_idx = __reduce_max_sync(0xffffffffu, last_idx);
  if (lane == 0) s_i[0][warp] = last_idx;
  __syncthreads();
  last_idx = s_i[0][0];
  for (int w = 1; w < GP_THREADS / 32; ++w) last_idx = max(last_idx, s_i[0][w]);
  __syncthreads();
  int ts_floor = r.ts_begin;                                // timestamps below this are forbidden
  if (last_idx >= 0) ts_floor = (int)srow[last_idx] + ((last_ts && !penult_ts) ? 0 : 1);
  const bool at_begin = nxt == begin_pos;
  const int text_hi = at_begin ? 0 : r.ts_begin;            // text tokens [text_lo, text_hi) are allowed
  const int text_lo = (last_ts && !penult_ts) ? r.eos : 0;
  const bool ts_allowed = !(last_ts && penult_ts);
  const int ts_hi = (at_begin && r.max_initial >= 0) ? min(vocab, r.ts_begin + r.max_initial + 1) : vocab;
  // ---- one scan
  const float* row = logits + (int64_t)b * ld;
  const bool use_begin_bias = bias_begin != nullptr && at_begin;
  float bt = -INFINITY, bs = -INFINITY, m_ts = -INFINITY, z_ts = 0.f;
  int bt_i = 0x7fffffff, bs_i = 0x7fffffff;
  for (int c = tid; c < vocab; c += GP_THREADS) {
    float v = row[c];
    if (bias_all) v += bias_all[c];
    if (use_begin_bias) v += bias_begin[c];
    if (c < r.ts_begin) {
      if (c == r.no_ts || c < text_lo || c >= text_hi) continue;
      if (v > bt) { bt = v; bt_i = c; }
    } else {
      if (!ts_allowed || c < ts_floor || c >= ts_hi) continue;
      if (v > bs) { bs = v; bs_i = c; }
      if (v > m_ts) { z_ts = z_ts * __expf(m_ts - v) + 1.f; m_ts = v; } else if (v > -INFINITY) { z_ts += __expf(v - m_ts); }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bt, o); const int oi = __shfl_xor_sync(0xffffffffu, bt_i, o);
    if (ov > bt || (ov == bt && oi < bt_i)) { bt = ov; bt_i = oi; }
    const float pv = __shfl_xor_sync(0xffffffffu, bs, o); const int pi = __shfl_xor_sync(0xffffffffu, bs_i, o);
    if (pv > bs || (pv == bs && pi < bs_i)) { bs = pv; bs_i = pi; }
    const float om = __shfl_xor_sync(0xffffffffu, m_ts, o), oz = __shfl_xor_sync(0xffffffffu, z_ts, o);
    const float mm = fmaxf(m_ts, om);
    if (mm > -INFINITY) { z_ts = z_ts * __expf(m_ts - mm) + oz * __expf(om - mm); m_ts = mm; }
  }
  if (lane == 0) { s_f[0][warp] = bt; s_i[1][warp] = bt_i; s_f[1][warp] = bs; s_i[2][warp] = bs_i; s_f[2][warp] = m_ts; s_f[3][warp] = z_ts; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < GP_THREADS / 32; ++w) {
      if (s_f[0][w] > bt || (s_f[0][w] == bt && s_i[1][w] < bt_i)) { bt = s_f[0][w]; bt_i = s_i[1][w]; }
      if (s_f[1][w] > bs || (s_f[1][w] == bs && s_i[2][w] < bs_i)) { bs = s_f[1][w]; bs_i = s_i[2][w]; }
      const float mm = fmaxf(m_ts, s_f[2][w]);
      if (mm > -INFINITY) { z_ts = z_ts * __expf(m_ts - mm) + s_f[3][w] * __expf(s_f[2][w] - mm); m_ts = mm; }
    }
    const float lse_ts = z_ts > 0.f ? m_ts + __logf(z_ts) : -INFINITY;
    int pick;
    if (lse_ts > bt) pick = bs_i;                           // the timestamps together outweigh every text token
    else pick = (bs > bt) ? bs_i : bt_i;                    // plain arg-max (text indices are lower: ties go to text)
    if (pick == 0x7fffffff) pick = (int)eos;                // everything masked (cannot happen with a sane config)
    seq[(int64_t)b * seq_ld + nxt] = pick;
    if ((int64_t)pick == eos) finished[b] = 1;
  }
}

// pos += 1; done_at = first position count at which every row had finished (0 while some row is still decoding)
__global__ void decode_advance_kernel(int* __restrict__ pos_dev, const int* __restrict__ finished, int B, int* __restrict__ done_at) {
  int all = 1;
  for (int i = threadIdx.x; i < B; i += 32) all &= finished[i] != 0;
  all = __all_sync(0xffffffffu, all);
  if (threadIdx.x == 0) {
    const int p = *pos_dev + 1;
    *pos_dev = p;
    if (all && *done_at == 0) *done_at = p + 1;    // sequences are complete up to and including index p
  }
}

}  // namespace dwb

using namespace dwb;

extern "C" int dwb_embed_decode(const int64_t* seq, int seq_ld, const int* pos_dev, const void* E, const void* P, int table_is_f32,
                                float* x, int B, int d, int vocab, void* stream) {
  DWB_CHECK_ARG(seq && pos_dev && E && P && x && B > 0 && d > 0 && seq_ld > 0, "dwb_embed_decode: bad args");
  if (table_is_f32)
    embed_decode_kernel<float><<<B, 256, 0, (cudaStream_t)stream>>>(seq, seq_ld, pos_dev, (const float*)E, (const float*)P, x, d, vocab);
  else
    embed_decode_kernel<bf16><<<B, 256, 0, (cudaStream_t)stream>>>(seq, seq_ld, pos_dev, (const bf16*)E, (const bf16*)P, x, d, vocab);
  DWB_LAUNCH_OK();
  return DWB_OK;
}

extern "C" int dwb_attention_decode(const void* q, int64_t ldq, const void* k_new, const void* v_new, int64_t ld_new, void* k_cache,
                                    void* v_cache, int64_t ld_cache, int cache_rows, void* o, int64_t ldo, int B, int H, int head_dim,
                                    int fixed_len, const int* pos_dev, float scale, void* stream) {
  DWB_CHECK_ARG(head_dim == 64, "dwb_attention_decode: head_dim %d unsupported (Whisper uses 64)", head_dim);
  DWB_CHECK_ARG(q && k_cache && v_cache && o && B > 0 && H > 0 && cache_rows > 0, "dwb_attention_decode: bad args");
  DWB_CHECK_ARG((k_new == nullptr) == (v_new == nullptr), "dwb_attention_decode: k_new and v_new go together");
  DWB_CHECK_ARG(k_new ? pos_dev != nullptr : (fixed_len > 0 && fixed_len <= cache_rows), "dwb_attention_decode: need pos_dev (append) or 0 < fixed_len <= cache_rows");
  DWB_CHECK_ARG((ldq % 8) == 0 && (ld_cache % 8) == 0 && (ldo % 2) == 0 && (k_new == nullptr || (ld_new % 8) == 0),
                "dwb_attention_decode: row pitches must keep 16 B alignment");
  const size_t smem = (size_t)cache_rows * sizeof(float);
  DWB_CHECK_ARG(smem <= 200 * 1024, "dwb_attention_decode: %d cache rows exceed the shared-memory score buffer", cache_rows);
  static size_t smem_set = 48 * 1024;
  if (smem > smem_set) {
    DWB_CUDA_OK(cudaFuncSetAttribute(attn_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  attn_decode_kernel<<<dim3(H, B), AD_THREADS, smem, (cudaStream_t)stream>>>(
      (const bf16*)q, ldq, (const bf16*)k_new, (const bf16*)v_new, ld_new, (bf16*)k_cache, (bf16*)v_cache, ld_cache, cache_rows, (bf16*)o, ldo,
      fixed_len, pos_dev, scale * 1.4426950408889634f);
  DWB_LAUNCH_OK();
  return DWB_OK;
}

extern "C" int dwb_gemm_skinny_bf16(const void* X, int64_t ldx, const void* W, int64_t ldw, void* C, int64_t ldc, int c_f32, int M, int N, int K,
                                    const float* bias, int act, void* stream) {
  DWB_CHECK_ARG(X && W && C, "dwb_gemm_skinny_bf16: null operand");
  DWB_CHECK_ARG(M > 0 && M <= 64 && (M % 16) == 0 && (N % 8) == 0 && (K % 256) == 0, "dwb_gemm_skinny_bf16: needs M in {16,32,48,64}, N %% 8 == 0, K %% 256 == 0 (M=%d N=%d K=%d)", M, N, K);
  DWB_CHECK_ARG((ldx % 8) == 0 && (ldw % 8) == 0 && (ldc % 2) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(C) & (c_f32 ? 7 : 3)) == 0, "dwb_gemm_skinny_bf16: X / W need 16 B aligned rows, C 4 B (8 B fp32) pairs");
  DWB_CHECK_ARG(act == 0 || act == 1, "dwb_gemm_skinny_bf16: unknown activation %d", act);
  const dim3 grid(N / 8);
  cudaStream_t st = (cudaStream_t)stream;
  switch (M / 16) {
    case 1: gemm_skinny_kernel<1><<<grid, 128, 0, st>>>((const bf16*)X, ldx, (const bf16*)W, ldw, bias, C, ldc, c_f32, N, K, act); break;
    case 2: gemm_skinny_kernel<2><<<grid, 128, 0, st>>>((const bf16*)X, ldx, (const bf16*)W, ldw, bias, C, ldc, c_f32, N, K, act); break;
    case 3: gemm_skinny_kernel<3><<<grid, 128, 0, st>>>((const bf16*)X, ldx, (const bf16*)W, ldw, bias, C, ldc, c_f32, N, K, act); break;
    default: gemm_skinny_kernel<4><<<grid, 128, 0, st>>>((const bf16*)X, ldx, (const bf16*)W, ldw, bias, C, ldc, c_f32, N, K, act); break;
  }
  DWB_LAUNCH_OK();
  return DWB_OK;
}

extern "C" int dwb_greedy_pick(const float* logits, int64_t ld, int vocab, const float* bias_all, const float* bias_begin, int begin_pos,
                               int64_t* seq, int seq_ld, int prompt_len, int* finished, int64_t eos, int64_t pad, const int* pos_dev, int B,
                               void* stream) {
  DWB_CHECK_ARG(logits && seq && finished && pos_dev && B > 0 && vocab > 0 && ld >= vocab && prompt_len >= 1 && seq_ld >= prompt_len,
                "dwb_greedy_pick: bad args");
  greedy_pick_kernel<<<B, GP_THREADS, 0, (cudaStream_t)stream>>>(logits, ld, vocab, bias_all, bias_begin, begin_pos, seq, seq_ld, prompt_len,
                                                                 finished, eos, pad, pos_dev);
  DWB_LAUNCH_OK();
  return DWB_OK;
}

extern "C" int dwb_greedy_pick_timestamps(const float* logits, int64_t ld, int vocab, const float* bias_all, const float* bias_begin,
                                         int begin_pos, int64_t* seq, int seq_ld, int prompt_len, int* finished, int64_t eos, int64_t pad,
                                         const int* pos_dev, int B, int timestamp_begin, int max_initial_timestamp_index, void* stream) {
  DWB_CHECK_ARG(logits && seq && finished && pos_dev && B > 0 && vocab > 0 && ld >= vocab && prompt_len >= 1 && seq_ld >= prompt_len,
                "dwb_greedy_pick_timestamps: bad args");
  DWB_CHECK_ARG(timestamp_begin > 1 && timestamp_begin <= vocab && eos >= 0 && eos < timestamp_begin,
                "dwb_greedy_pick_timestamps: timestamp_begin %d / eos %lld inconsistent with vocab %d", timestamp_begin, (long long)eos, vocab);
  TimestampRules r;
  r.ts_begin = timestamp_begin; r.no_ts = timestamp_begin - 1; r.eos = (int)eos; r.max_initial = max_initial_timestamp_index;
  greedy_pick_ts_kernel<<<B, GP_THREADS, 0, (cudaStream_t)stream>>>(logits, ld, vocab, bias_all, bias_begin, begin_pos, seq, seq_ld, prompt_len,
                                                                    finished, eos, pad, pos_dev, r);
  DWB_LAUNCH_OK();
  return DWB_OK;
}

extern "C" int dwb_decode_advance(int* pos_dev, const int* finished, int B, int* done_at, void* stream) {
  DWB_CHECK_ARG(pos_dev && finished && done_at && B > 0, "dwb_decode_advance: bad args");
  decode_advance_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pos_dev, finished, B, done_at);
  DWB_LAUNCH_OK();
  return DWB_OK;
}
