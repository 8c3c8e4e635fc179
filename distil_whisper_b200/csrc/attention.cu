// Multi-head attention, head_dim 64 (every Whisper size), bf16 in / fp32 softmax / bf16 out.
// Generic flash-style kernels built on mma.sync (m16n8k16): forward (+LSE), backward preprocess, backward.
// Used for decoder self-attention (causal, T<=448), decoder cross-attention (T x 1500) and their gradients;
// the big non-causal encoder forward has a tcgen05 version in attention_tcgen05.cu.
//
// Replaces F.scaled_dot_product_attention as reached through HF:integrations/sdpa_attention.py:40-104 from
// HF:models/whisper/modeling_whisper.py:342-352.  Semantics kept: q is pre-scaled by head_dim^-0.5 in the reference
// (:310) and sdpa runs with scaling=1.0; here the same factor is folded into the exponent (`scale`), which is
// bit-identical for a power-of-two factor.  is_causal == lower-triangular mask aligned at position 0 (:77).
//
// Layout: Q rows live in a [B*Sq, ldq] matrix, head h in columns [h*64, h*64+64) (so the fused QKV GEMM output is
// consumed in place); same for K, V ([B*Sk, ldk/ldv]) and O ([B*Sq, ldo]).  LSE is [B, H, Sq] fp32.
#include "common.cuh"

namespace dwb {

constexpr int HD = 64;
constexpr int ATT_BR = 64;   // query rows per CTA (4 warps x 16)
constexpr int ATT_BC = 64;   // keys per inner tile

__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g, bool valid) {
  const int sz = valid ? 16 : 0;   // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(g), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// smem tile: 64 rows x 64 bf16 (128 B rows), 16 B chunks XOR-swizzled with (row & 7)
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int col /*multiple of 8*/) {
  return base + row * 128 + ((((col >> 3) ^ (row & 7)) & 7) << 4);
}
// async-load a [64 x 64] bf16 tile: rows r0.. of a matrix with `ld` elements per row, rows >= nrows zero filled
__device__ __forceinline__ void load_tile_async(uint32_t sbase, const bf16* g, int64_t ld, int r0, int nrows, int tid,
                                                int nthreads) {
  for (int i = tid; i < 64 * 8; i += nthreads) {
    const int r = i >> 3, c = (i & 7) << 3;
    const bool ok = (r0 + r) < nrows;
    const bf16* src = g + (int64_t)(ok ? (r0 + r) : 0) * ld + c;
    cp_async16(tile_addr(sbase, r, c), src, ok);
  }
}

struct AttnParams {
  const bf16 *q, *k, *v;
  bf16* o;
  float* lse;
  int64_t ldq, ldk, ldv, ldo;
  int B, H, Sq, Sk;
  int causal;
  float scale;
};

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_fwd_kernel(const AttnParams p) {
  __shared__ __align__(128) uint8_t smem[(64 * 128) * 5];   // Q, K0, K1, V0, V1 : 40 KB
  const uint32_t sQ = smem_u32(smem), sK = sQ + 8192, sV = sQ + 3 * 8192;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y, b = bh / p.H, h = bh % p.H;
  const int q0 = blockIdx.x * ATT_BR;
  const bf16* Q = p.q + (int64_t)b * p.Sq * p.ldq + h * HD;
  const bf16* K = p.k + (int64_t)b * p.Sk * p.ldk + h * HD;
  const bf16* V = p.v + (int64_t)b * p.Sk * p.ldv + h * HD;

  int kv_end = p.Sk;
  if (p.causal) kv_end = min(p.Sk, q0 + ATT_BR);
  const int n_tiles = ceil_div(kv_end, ATT_BC);

  load_tile_async(sQ, Q, p.ldq, q0, p.Sq, tid, 128);
  load_tile_async(sK, K, p.ldk, 0, p.Sk, tid, 128);
  load_tile_async(sV, V, p.ldv, 0, p.Sk, tid, 128);
  cp_async_commit();

  float o_acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const float sl2 = p.scale * 1.4426950408889634f;
  uint32_t qf[4][4];
  const int row_a = q0 + warp * 16 + (lane >> 2);   // this thread's rows: row_a and row_a + 8

  for (int t = 0; t < n_tiles; ++t) {
    const int st = t & 1;
    if (t + 1 < n_tiles) {
      load_tile_async(sK + (st ^ 1) * 8192, K, p.ldk, (t + 1) * ATT_BC, p.Sk, tid, 128);
      load_tile_async(sV + (st ^ 1) * 8192, V, p.ldv, (t + 1) * ATT_BC, p.Sk, tid, 128);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (t == 0) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        ldsm_x4(qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3],
                tile_addr(sQ, warp * 16 + (lane & 15), kk * 16 + (lane >> 4) * 8));
    }
    // S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    const uint32_t kb = sK + st * 8192;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(b0, b1, b2, b3, tile_addr(kb, np * 16 + (lane & 7) + (lane >> 4) * 8, kk * 16 + ((lane >> 3) & 1) * 8));
        mma_bf16(s[2 * np], qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], b0, b1);
        mma_bf16(s[2 * np + 1], qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], b2, b3);
      }
    }
    // mask (key tail / causal) + online softmax
    const int c_base = t * ATT_BC + (lane & 3) * 2;
    const bool need_mask = (t * ATT_BC + ATT_BC > p.Sk) || (p.causal && (t * ATT_BC + ATT_BC > q0 + warp * 16));
    float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (need_mask) {
          const int col = c_base + i * 8 + (e & 1);
          const int row = row_a + (e >> 1) * 8;
          if (col >= p.Sk || (p.causal && col > row)) s[i][e] = -INFINITY;
        }
        mx[e >> 1] = fmaxf(mx[e >> 1], s[i][e]);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], msc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float mnew = mx[r];
      msc[r] = (mnew == -INFINITY) ? 0.f : mnew * sl2;
      corr[r] = (m_run[r] == -INFINITY) ? 0.f : exp2f(m_run[r] * sl2 - msc[r]);
      m_run[r] = mnew;
      l_run[r] *= corr[r];
    }
    float ls[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = exp2f(s[i][e] * sl2 - msc[e >> 1]);
        s[i][e] = pv;
        ls[e >> 1] += pv;
      }
      o_acc[i][0] *= corr[0]; o_acc[i][1] *= corr[0];
      o_acc[i][2] *= corr[1]; o_acc[i][3] *= corr[1];
    }
    l_run[0] += ls[0];
    l_run[1] += ls[1];
    // O += P V
    const uint32_t vb = sV + st * 8192;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {   // 16 keys per step
      const uint32_t a0 = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]), a1 = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      const uint32_t a2 = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]), a3 = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // 16 head-dim columns per step
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(b0, b1, b2, b3, tile_addr(vb, kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, np * 16 + (lane >> 4) * 8));
        mma_bf16(o_acc[2 * np], a0, a1, a2, a3, b0, b1);
        mma_bf16(o_acc[2 * np + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    __syncthreads();
  }
  // finalize
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = row_a + r * 8;
    if (row < p.Sq) {
      const float inv = l_run[r] > 0.f ? 1.f / l_run[r] : 0.f;
      bf16* orow = p.o + ((int64_t)b * p.Sq + row) * p.ldo + h * HD + (lane & 3) * 2;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t pk = pack_bf16x2(o_acc[i][2 * r] * inv, o_acc[i][2 * r + 1] * inv);
        *reinterpret_cast<uint32_t*>(orow + i * 8) = pk;
      }
      if (p.lse != nullptr && (lane & 3) == 0)
        p.lse[((int64_t)b * p.H + h) * p.Sq + row] =
            (l_run[r] > 0.f) ? (m_run[r] * p.scale + logf(l_run[r])) : -INFINITY;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward preprocess: delta[b,h,i] = sum_d dO[i,d] * O[i,d]
__global__ void attn_bwd_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout, float* __restrict__ delta,
                                      int64_t ldo, int64_t lddo, int B, int H, int Sq) {
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int total = B * H * Sq;
  if (warp_global >= total) return;
  const int i = warp_global % Sq, bh = warp_global / Sq, h = bh % H, b = bh / H;
  const bf16* po = o + ((int64_t)b * Sq + i) * ldo + h * HD + lane * 2;
  const bf16* pd = dout + ((int64_t)b * Sq + i) * lddo + h * HD + lane * 2;
  const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(po));
  const float2 g = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(pd));
  const float s = warp_sum(a.x * g.x + a.y * g.y);
  if (lane == 0) delta[warp_global] = s;
}

struct AttnBwdParams {
  const bf16 *q, *k, *v, *dout;
  const float *lse, *delta;
  float* dq_acc;           // fp32 [B*Sq, H*64] accumulated with atomics (zeroed by the caller)
  bf16 *dk, *dv;
  int64_t ldq, ldk, ldv, lddo, lddk, lddv;
  int B, H, Sq, Sk;
  int causal;
  float scale;
};

// one CTA = one (kv tile of 64 keys, b, h); warp w owns kv rows [16w, 16w+16) of the tile
__global__ void __launch_bounds__(128) attn_bwd_kernel(const AttnBwdParams p) {
  extern __shared__ __align__(128) uint8_t smem_dyn[];
  const uint32_t sK = smem_u32(smem_dyn), sV = sK + 8192, sQ = sK + 2 * 8192 /*2 stages*/, sDO = sK + 4 * 8192 /*2 stages*/,
                 sDS = sK + 6 * 8192;
  float* sLse = reinterpret_cast<float*>(smem_dyn + 7 * 8192);   // [2][64]
  float* sDelta = sLse + 128;                                    // [2][64]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y, b = bh / p.H, h = bh % p.H;
  const int k0 = blockIdx.x * ATT_BC;
  const bf16* Q = p.q + (int64_t)b * p.Sq * p.ldq + h * HD;
  const bf16* K = p.k + (int64_t)b * p.Sk * p.ldk + h * HD;
  const bf16* V = p.v + (int64_t)b * p.Sk * p.ldv + h * HD;
  const bf16* DO = p.dout + (int64_t)b * p.Sq * p.lddo + h * HD;
  const float* LSE = p.lse + ((int64_t)b * p.H + h) * p.Sq;
  const float* DEL = p.delta + ((int64_t)b * p.H + h) * p.Sq;

  int qt_begin = 0;
  if (p.causal) qt_begin = k0 / ATT_BR;          // queries before k0 never see this kv tile
  const int qt_end = ceil_div(p.Sq, ATT_BR);

  load_tile_async(sK, K, p.ldk, k0, p.Sk, tid, 128);
  load_tile_async(sV, V, p.ldv, k0, p.Sk, tid, 128);
  if (qt_begin < qt_end) {
    load_tile_async(sQ, Q, p.ldq, qt_begin * ATT_BR, p.Sq, tid, 128);
    load_tile_async(sDO, DO, p.lddo, qt_begin * ATT_BR, p.Sq, tid, 128);
  }
  cp_async_commit();

  float dk_acc[8][4], dv_acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) dk_acc[i][e] = dv_acc[i][e] = 0.f;
  uint32_t kf[4][4], vf[4][4];
  const int kv_row_a = k0 + warp * 16 + (lane >> 2);

  for (int qt = qt_begin; qt < qt_end; ++qt) {
    const int st = (qt - qt_begin) & 1;
    const int q0 = qt * ATT_BR;
    if (tid < 64) {
      const int qi = q0 + tid;
      sLse[st * 64 + tid] = qi < p.Sq ? LSE[qi] : INFINITY;    // +inf -> P = 0 for padded query rows
      sDelta[st * 64 + tid] = qi < p.Sq ? DEL[qi] : 0.f;
    }
    if (qt + 1 < qt_end) {
      load_tile_async(sQ + (st ^ 1) * 8192, Q, p.ldq, (qt + 1) * ATT_BR, p.Sq, tid, 128);
      load_tile_async(sDO + (st ^ 1) * 8192, DO, p.lddo, (qt + 1) * ATT_BR, p.Sq, tid, 128);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (qt == qt_begin) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        ldsm_x4(kf[kk][0], kf[kk][1], kf[kk][2], kf[kk][3], tile_addr(sK, warp * 16 + (lane & 15), kk * 16 + (lane >> 4) * 8));
        ldsm_x4(vf[kk][0], vf[kk][1], vf[kk][2], vf[kk][3], tile_addr(sV, warp * 16 + (lane & 15), kk * 16 + (lane >> 4) * 8));
      }
    }
    const uint32_t qb = sQ + st * 8192, dob = sDO + st * 8192;
    // S^T = K Q^T, dP^T = V dO^T   (16 kv rows x 64 query cols per warp)
    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[i][e] = dp[i][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        const int r = np * 16 + (lane & 7) + (lane >> 4) * 8, c = kk * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4(b0, b1, b2, b3, tile_addr(qb, r, c));
        mma_bf16(s[2 * np], kf[kk][0], kf[kk][1], kf[kk][2], kf[kk][3], b0, b1);
        mma_bf16(s[2 * np + 1], kf[kk][0], kf[kk][1], kf[kk][2], kf[kk][3], b2, b3);
        ldsm_x4(b0, b1, b2, b3, tile_addr(dob, r, c));
        mma_bf16(dp[2 * np], vf[kk][0], vf[kk][1], vf[kk][2], vf[kk][3], b0, b1);
        mma_bf16(dp[2 * np + 1], vf[kk][0], vf[kk][1], vf[kk][2], vf[kk][3], b2, b3);
      }
    }
    // P^T = exp(S^T * scale - lse[q]);  dS^T = P^T * (dP^T - delta[q]) * scale
    uint32_t pT[8][2], dsT[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float pv[4], dsv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int qc = i * 8 + (lane & 3) * 2 + (e & 1);        // query column inside the tile
        const int kv = kv_row_a + (e >> 1) * 8;
        const float l = sLse[st * 64 + qc];
        float pe = __expf(s[i][e] * p.scale - l);
        if (kv >= p.Sk || (p.causal && kv > q0 + qc)) pe = 0.f;
        pv[e] = pe;
        dsv[e] = pe * (dp[i][e] - sDelta[st * 64 + qc]) * p.scale;
      }
      pT[i][0] = pack_bf16x2(pv[0], pv[1]);   pT[i][1] = pack_bf16x2(pv[2], pv[3]);
      dsT[i][0] = pack_bf16x2(dsv[0], dsv[1]); dsT[i][1] = pack_bf16x2(dsv[2], dsv[3]);
      // stash dS^T (bf16) for the dQ product: tile [64 kv][64 q]
      const int r0 = warp * 16 + (lane >> 2), cc = i * 8 + (lane & 3) * 2;
      const uint32_t a0 = tile_addr(sDS, r0, i * 8) + (lane & 3) * 4;
      const uint32_t a1 = tile_addr(sDS, r0 + 8, i * 8) + (lane & 3) * 4;
      (void)cc;
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(a0), "r"(dsT[i][0]) : "memory");
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(a1), "r"(dsT[i][1]) : "memory");
    }
    // dV += P^T dO ; dK += dS^T Q     (k dim = the 64 queries of this tile)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t pa0 = pT[2 * kk][0], pa1 = pT[2 * kk][1], pa2 = pT[2 * kk + 1][0], pa3 = pT[2 * kk + 1][1];
      const uint32_t da0 = dsT[2 * kk][0], da1 = dsT[2 * kk][1], da2 = dsT[2 * kk + 1][0], da3 = dsT[2 * kk + 1][1];
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        const int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, c = np * 16 + (lane >> 4) * 8;
        ldsm_x4_t(b0, b1, b2, b3, tile_addr(dob, r, c));
        mma_bf16(dv_acc[2 * np], pa0, pa1, pa2, pa3, b0, b1);
        mma_bf16(dv_acc[2 * np + 1], pa0, pa1, pa2, pa3, b2, b3);
        ldsm_x4_t(b0, b1, b2, b3, tile_addr(qb, r, c));
        mma_bf16(dk_acc[2 * np], da0, da1, da2, da3, b0, b1);
        mma_bf16(dk_acc[2 * np + 1], da0, da1, da2, da3, b2, b3);
      }
    }
    __syncthreads();   // dS^T tile complete
    // dQ[16 queries of this warp, :] += dS[q, kv] K[kv, :]   -> fp32 atomics
    {
      float dq[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {          // 16 kv per step
        uint32_t a0, a1, a2, a3;
        // A = dS (rows q, cols kv) read transposed from the [kv][q] tile
        ldsm_x4_t(a0, a1, a2, a3, tile_addr(sDS, kk * 16 + (lane & 7) + (lane >> 4) * 8, warp * 16 + ((lane >> 3) & 1) * 8));
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(b0, b1, b2, b3, tile_addr(sK, kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, np * 16 + (lane >> 4) * 8));
          mma_bf16(dq[2 * np], a0, a1, a2, a3, b0, b1);
          mma_bf16(dq[2 * np + 1], a0, a1, a2, a3, b2, b3);
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int qi = q0 + warp * 16 + (lane >> 2) + r * 8;
        if (qi < p.Sq) {
          float* dst = p.dq_acc + ((int64_t)b * p.Sq + qi) * (p.H * HD) + h * HD + (lane & 3) * 2;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            atomicAdd(dst + i * 8, dq[i][2 * r]);
            atomicAdd(dst + i * 8 + 1, dq[i][2 * r + 1]);
          }
        }
      }
    }
    __syncthreads();   // before the next iteration overwrites sDS / the other Q,dO stage
  }
  cp_async_wait<0>();
  // write dK, dV
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int kv = kv_row_a + r * 8;
    if (kv < p.Sk) {
      bf16* dkr = p.dk + ((int64_t)b * p.Sk + kv) * p.lddk + h * HD + (lane & 3) * 2;
      bf16* dvr = p.dv + ((int64_t)b * p.Sk + kv) * p.lddv + h * HD + (lane & 3) * 2;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        *reinterpret_cast<uint32_t*>(dkr + i * 8) = pack_bf16x2(dk_acc[i][2 * r], dk_acc[i][2 * r + 1]);
        *reinterpret_cast<uint32_t*>(dvr + i * 8) = pack_bf16x2(dv_acc[i][2 * r], dv_acc[i][2 * r + 1]);
      }
    }
  }
}

}  // namespace dwb

using namespace dwb;

extern "C" int dwb_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                                 int64_t ldo, float* lse, int B, int H, int Sq, int Sk, int head_dim, int causal,
                                 float scale, void* stream) {
  DWB_CHECK_ARG(head_dim == HD, "dwb_attention_fwd: head_dim %d unsupported (Whisper uses 64)", head_dim);
  DWB_CHECK_ARG(q && k && v && o, "dwb_attention_fwd: null operand");
  DWB_CHECK_ARG(B > 0 && H > 0 && Sq > 0 && Sk > 0, "dwb_attention_fwd: bad shape");
  DWB_CHECK_ARG((ldq % 8) == 0 && (ldk % 8) == 0 && (ldv % 8) == 0 && (ldo % 2) == 0, "dwb_attention_fwd: row pitch alignment");
  AttnParams p{(const bf16*)q, (const bf16*)k, (const bf16*)v, (bf16*)o, lse, ldq, ldk, ldv, ldo, B, H, Sq, Sk, causal, scale};
  dim3 grid(ceil_div(Sq, ATT_BR), B * H);
  attn_fwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(p);
  DWB_LAUNCH_OK();
  return DWB_OK;
}

extern "C" int dwb_attention_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                 const void* o, int64_t ldo, const void* dout, int64_t lddo, const float* lse,
                                 float* delta_ws, float* dq_acc, void* dk, int64_t lddk, void* dv, int64_t lddv, int B, int H,
                                 int Sq, int Sk, int head_dim, int causal, float scale, void* stream) {
  DWB_CHECK_ARG(head_dim == HD, "dwb_attention_bwd: head_dim %d unsupported", head_dim);
  DWB_CHECK_ARG(q && k && v && o && dout && lse && delta_ws && dq_acc && dk && dv, "dwb_attention_bwd: null operand");
  cudaStream_t st = (cudaStream_t)stream;
  const int total = B * H * Sq;
  attn_bwd_delta_kernel<<<ceil_div(total, 8), 256, 0, st>>>((const bf16*)o, (const bf16*)dout, delta_ws, ldo, lddo, B, H, Sq);
  DWB_LAUNCH_OK();
  DWB_CUDA_OK(cudaMemsetAsync(dq_acc, 0, (size_t)B * Sq * H * HD * sizeof(float), st));
  AttnBwdParams p{(const bf16*)q, (const bf16*)k, (const bf16*)v, (const bf16*)dout, lse, delta_ws, dq_acc, (bf16*)dk, (bf16*)dv,
                  ldq, ldk, ldv, lddo, lddk, lddv, B, H, Sq, Sk, causal, scale};
  const int smem = 7 * 8192 + 4 * 64 * 4;
  static bool attr = false;
  if (!attr) {
    DWB_CUDA_OK(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  dim3 grid(ceil_div(Sk, ATT_BC), B * H);
  attn_bwd_kernel<<<grid, 128, smem, st>>>(p);
  DWB_LAUNCH_OK();
  return DWB_OK;
}
