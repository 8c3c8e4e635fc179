// Non-causal attention forward on tcgen05 / TMEM (head_dim 64): the Whisper encoder self-attention, S = 1500.
//   per CTA: one (batch, head, 128-query tile); loops over 128-key tiles
//   S = Q K^T      tcgen05.mma 128x128x64  (Q, K tiles: TMA, 128B swizzle, K-major)        -> TMEM cols [0,128)
//   softmax        one thread per query row (tcgen05.ld 32x32b: lane == row, no shuffles), online max/sum in fp32,
//                  P (bf16 pairs) written back to TMEM (tcgen05.st) and consumed from there as the A operand of P V
//   O += P V       tcgen05.mma 128x64x128  (V tile is the MN-major B operand: no transpose)  -> TMEM cols [128,192),
//                  accumulated in TMEM across key tiles; rescaled (tcgen05.ld/st) only when a row maximum grows by > 2^8
//                  ("lazy rescale"), normalised by the row sum at the end, TMA store
// Two CTAs are resident per SM (112 KB smem, 256 TMEM columns each), so one CTA's exp2-bound softmax overlaps the
// other's MMAs; inside a CTA, QK^T of tile j+1 is issued as soon as tile j's scores have left TMEM.
//
// Replaces F.scaled_dot_product_attention (HF:integrations/sdpa_attention.py:40-104) for
// HF:models/whisper/modeling_whisper.py:342-352 in the encoder (is_causal False), same contract as dwb_attention_fwd.
#include <stdlib.h>

#include "common.cuh"

namespace dwb {

constexpr int TA_BQ = 128, TA_BK = 128, TA_HD = 64;
constexpr int TA_THREADS = 256;   // warps 0-3 softmax warpgroup; warp 4 TMA, warp 5 MMA, warps 6-7 idle (complete the 2nd warpgroup)
constexpr int TA_TILE_BYTES = 128 * 128;            // 128 rows x 128 B
constexpr int TA_KV_STAGES = 3;
constexpr int TA_TILES_BYTES = TA_TILE_BYTES /*Q, reused for the output tile*/ + TA_KV_STAGES * 2 * TA_TILE_BYTES /*K,V*/;
constexpr int TA_BAR_BYTES = 112;
// two CTAs per SM: 2 * (dyn + 1 KB system reserve) <= 228 KB  ->  dyn <= 115712; the slack absorbs the 1 KB round-up
constexpr int TA_SMEM = 115712;
static_assert(TA_TILES_BYTES + TA_BAR_BYTES + 896 <= TA_SMEM, "smem budget");
constexpr int TA_TMEM_COLS = 256;
constexpr int TA_DEFAULT_VARIANT = 1;   // np8 + 10 * mode (see dwb_attention_fwd_tc)

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// 2^x for a pair of scores on the FMA / ALU pipes instead of the MUFU (the XU pipe, 16 exp2 per clock per SM, is what bounds
// this kernel at head_dim 64): Cody-Waite split x = n + f with n = round(x) taken from the float's own mantissa
// (x + 1.5 * 2^23), 2^f by a degree-3 minimax polynomial on [-0.5, 0.5] (max relative error 7.5e-5, fifty times below the bf16
// rounding P receives next), and n added into the exponent field.  x is clamped to [-126, 126]: a masked score (-inf) becomes
// 2^-126 (nothing), an overflowing one 2^126, which the tile-sum test of the lazy rescale still catches.
__device__ __forceinline__ float2 poly_exp2_pair(float2 x) {
  x.x = fminf(fmaxf(x.x, -126.f), 126.f);
  x.y = fminf(fmaxf(x.y, -126.f), 126.f);
  const float2 magic = make_float2(12582912.f, 12582912.f), neg_magic = make_float2(-12582912.f, -12582912.f);
  const float2 t = __fadd2_rn(x, magic);
  const float2 n = __fadd2_rn(t, neg_magic);
  const float2 f = __ffma2_rn(n, make_float2(-1.f, -1.f), x);
  float2 q = __ffma2_rn(make_float2(0.055171459913253784f, 0.055171459913253784f), f, make_float2(0.2426108568906784f, 0.2426108568906784f));
  q = __ffma2_rn(q, f, make_float2(0.6932609677314758f, 0.6932609677314758f));
  q = __ffma2_rn(q, f, make_float2(0.9999281167984009f, 0.9999281167984009f));
  return make_float2(__int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23)),
                     __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23)));
}

struct TcAttnParams {
  int H, Sq, Sk, causal;
  int reverse;          // CTAs take (batch, head, query tile) from the last to the first (dwb_set_row_walk)
  float scale_log2;     // scale * log2(e)
  float scale;
  float* lse;           // [B, H, Sq] or null
};

// NP8: of every 8 score pairs, NP8 take the polynomial exp2 (spread evenly through the unrolled loop so that FMA-pipe chains
// fill the issue slots between MUFU issues); 0 = all MUFU.
// MODE 1 (DEFER): the wait for "P V of the previous tile has consumed P" is taken after the first 64 exponentials of the tile
// are in registers instead of before the first one, so that the previous tile's P V (and its completion signalling) runs under them.
// MODE 2 (PIPE): software-pipelined softmax loop -- see the branch below.
template <int NP8, int MODE>
__global__ void __launch_bounds__(TA_THREADS, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                   const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o,
                   const TcAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (threadIdx.x == 0 && (smem - smem_raw) + TA_TILES_BYTES + TA_BAR_BYTES > TA_SMEM) {
    printf("dwb: attention smem window misaligned by %d B\n", (int)(smem - smem_raw));
    __trap();
  }
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + TA_TILE_BYTES;                  // stage s: K at s*32K, V at s*32K + 16K
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + TA_KV_STAGES * 2 * TA_TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;      // [3]
  uint64_t* kv_empty = bars + 4;     // [3]
  uint64_t* s_full = bars + 7;
  uint64_t* s_empty = bars + 8;
  uint64_t* p_full = bars + 9;
  uint64_t* p_empty = bars + 10;
  uint64_t* o_full = bars + 11;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bx = p.reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x, by = p.reverse ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
  const int q0 = bx * TA_BQ;
  const int b = by / p.H, h = by % p.H;
  // causal: keys beyond the last query of this tile are never attended to
  const int n_kv = ceil_div(p.causal ? min(p.Sk, q0 + TA_BQ) : p.Sk, TA_BK);

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v); tma_prefetch_desc(&tmap_o);
    mbar_init(q_full, 1);
    for (int i = 0; i < TA_KV_STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(s_full, 1); mbar_init(s_empty, 128);
    mbar_init(p_full, 128); mbar_init(p_empty, 1);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_ptr, TA_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_s = tmem_base;            // 128 columns: fp32 scores of the current key tile
  const uint32_t tmem_p = tmem_base + 128;      //  64 columns: bf16 probabilities, 2 keys per 32-bit column (A operand of P V)
  const uint32_t tmem_o = tmem_base + 192;      //  64 columns: fp32 output accumulator

  // register re-distribution between the two warpgroups (launch: 128/thread for 2 CTAs/SM): 208 for softmax, 48 for the rest
  if (warp >= 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
  }
  if (warp == 4) {
    // ===================================== TMA producer (warp-uniform loop, elected lane issues) ==========
    {
      const bool leader = elect_one();
      if (leader) {
        mbar_expect_tx(q_full, TA_TILE_BYTES);
        tma_load_3d(&tmap_q, q_full, sQ, h * TA_HD, q0, b);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % TA_KV_STAGES;
        mbar_wait(&kv_empty[st], ((j / TA_KV_STAGES) & 1) ^ 1);
        if (leader) {
          mbar_expect_tx(&kv_full[st], 2 * TA_TILE_BYTES);
          tma_load_3d(&tmap_k, &kv_full[st], sKV + st * 2 * TA_TILE_BYTES, h * TA_HD, j * TA_BK, b);
          tma_load_3d(&tmap_v, &kv_full[st], sKV + st * 2 * TA_TILE_BYTES + TA_TILE_BYTES, h * TA_HD, j * TA_BK, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 5) {
    // ===================================== MMA issuer (warp-uniform loop, elected lane issues) ============
    // issue order: QK(0), then per tile j: QK(j+1) (as soon as tile j's scores have left TMEM), PV(j) (once P_j is in TMEM)
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);
      mbar_wait(q_full, 0);
      const uint64_t dq = umma_desc_sw128(smem_u32(sQ), 16, 1024);
      auto issue_qk = [&](int j) {
        const int st = j % TA_KV_STAGES;
        mbar_wait(&kv_full[st], (j / TA_KV_STAGES) & 1);
        mbar_wait(s_empty, (j & 1) ^ 1);
        tc_fence_after();
        if (leader) {
          const uint64_t dk = umma_desc_sw128(smem_u32(sKV + st * 2 * TA_TILE_BYTES), 16, 1024);
#pragma unroll
          for (int k = 0; k < TA_HD / 16; ++k) tc_mma_ss(tmem_s, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_qk, k > 0 ? 1u : 0u);
          tc_commit(s_full);
        }
        __syncwarp();
      };
      issue_qk(0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % TA_KV_STAGES;
        if (j + 1 < n_kv) issue_qk(j + 1);
        mbar_wait(p_full, j & 1);                  // P_j is in TMEM (and any rescale of O finished)
        tc_fence_after();
        if (leader) {
          const uint64_t dv = umma_desc_sw128(smem_u32(sKV + st * 2 * TA_TILE_BYTES + TA_TILE_BYTES), TA_TILE_BYTES, 1024);
#pragma unroll
          for (int k = 0; k < TA_BK / 16; ++k)       // A = P from TMEM: 16 keys = 8 columns per step; O accumulates in TMEM
            tc_mma_ts(tmem_o, tmem_p + 8 * k, dv + (uint64_t)(k * 128), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
          tc_commit(o_full);
          tc_commit(&kv_empty[st]);
          tc_commit(p_empty);
        }
        __syncwarp();
      }
    }
  } else if (warp < 4) {
    // ===================================== softmax / output =================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    // One thread per query row.  O lives in TMEM and is only rescaled when the row maximum has grown by more than 2^8
    // relative to the reference maximum m_ref the accumulators are expressed in (bf16 P and fp32 sums absorb the 2^8).
    const int row = warp * 32 + lane;                       // query row of this thread == TMEM lane
    const uint32_t t_s = tmem_s + ((uint32_t)(warp * 32) << 16);
    const uint32_t t_o = tmem_o + ((uint32_t)(warp * 32) << 16);
    const uint32_t t_p = tmem_p + ((uint32_t)(warp * 32) << 16);
    const int sw = row & 7;
    float m_ref = -INFINITY, l_run = 0.f;
    constexpr bool DEFER = MODE == 1;

    if constexpr (MODE == 2) {
      // Software-pipelined variant.  The scores of tile j+1 are pulled from TMEM 32 columns at a time INTO THE REGISTERS THE
      // exponentials of tile j have just released, so the TMEM read (64 B/clk per sub-partition: 256 cycles per tile) runs under
      // the MUFU work instead of in front of it, and "scores are in registers" (-> Q K^T of tile j+2 may be issued) is signalled
      // one tile earlier.  The row maximum of every tile is taken up front (FMNMX3 tree, ~45 instructions on the otherwise idle
      // ALU pipe): exponentials are then bounded by 2^8 relative to the reference maximum, O is rescaled only when a row's
      // maximum grows by more than 2^8 ("lazy rescale"), and no tile is ever exponentiated twice.
      uint32_t v[128];
      mbar_wait(s_full, 0);
      tc_fence_after();
      tmem_ld_32x32(t_s, v);
      tmem_ld_32x32(t_s + 32, v + 32);
      tmem_ld_32x32(t_s + 64, v + 64);
      tmem_ld_32x32(t_s + 96, v + 96);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_empty);
      const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
      for (int j = 0; j < n_kv; ++j) {
        const int kbase = j * TA_BK;
        const bool tail = kbase + TA_BK > p.Sk;
        const bool diag = p.causal && (kbase + TA_BK - 1 > q0);
        if (tail || diag) {
          const int lim = diag ? min(p.Sk, q0 + row + 1) : p.Sk;
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (kbase + i >= lim) v[i] = __float_as_uint(-INFINITY);
        }
        float mxp[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) mxp[c] = __uint_as_float(v[c]);
#pragma unroll
        for (int i = 8; i < 128; i += 8) {
#pragma unroll
          for (int c = 0; c < 8; ++c) mxp[c] = fmaxf(mxp[c], __uint_as_float(v[i + c]));
        }
        const float tmax = fmaxf(fmaxf(fmaxf(mxp[0], mxp[1]), fmaxf(mxp[2], mxp[3])), fmaxf(fmaxf(mxp[4], mxp[5]), fmaxf(mxp[6], mxp[7])));
        bool p_waited = false;
        if (j == 0) {
          m_ref = tmax;
        } else if (__any_sync(0xffffffffu, (tmax - m_ref) * p.scale_log2 > 8.0f)) {
          // rare, warp-uniform: bring O (TMEM) and l to the new reference maximum; rows that do not need it keep theirs (f = 1)
          mbar_wait(p_empty, (j & 1) ^ 1);                  // P V of the previous tile has completed: O is stable
          tc_fence_after();
          p_waited = true;
          const float mx = fmaxf(m_ref, tmax);
          const float f = fast_exp2((m_ref - mx) * p.scale_log2);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld_32x32(t_o + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st_32x32(t_o + c * 32, o);
          }
          tmem_st_wait();
          l_run *= f;
          m_ref = mx;
        }
        const float msc = m_ref * p.scale_log2;
        const float2 nm2 = make_float2(-msc, -msc);
        float2 ls[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        const bool more = j + 1 < n_kv;
        auto exp_keys = [&](const int base, const int n, uint32_t* pk) {          // keys [base, base + n) -> n / 2 packed columns
#pragma unroll
          for (int i = 0; i < n; i += 2) {
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(v[base + i]), __uint_as_float(v[base + i + 1])), sc2, nm2);
            float2 e;
            if (((((base + i) >> 1) * NP8) & 7) < NP8) {
              e = poly_exp2_pair(x);
            } else {
              e.x = fast_exp2(x.x);
              e.y = fast_exp2(x.y);
            }
            ls[(i >> 1) & 3] = __fadd2_rn(ls[(i >> 1) & 3], e);
            pk[i >> 1] = pack_bf16x2(e.x, e.y);
          }
        };
        {
          // first half of the tile: its 64 exponentials run while P V of the previous tile and Q K^T of the next one complete
          uint32_t pk[32];
          exp_keys(0, 64, pk);
          if (!p_waited) {
            mbar_wait(p_empty, (j & 1) ^ 1);                // P V of the previous tile has consumed P
            tc_fence_after();
          }
          tmem_st_32x32(t_p, pk);
          if (more) {
            mbar_wait(s_full, (j + 1) & 1);                 // Q K^T of the next tile (issued when this tile's scores left TMEM)
            tc_fence_after();
            tmem_ld_32x32(t_s, v);                          // next tile's scores into the registers this half has released
            tmem_ld_32x32(t_s + 32, v + 32);
          }
        }
#pragma unroll
        for (int c = 2; c < 4; ++c) {                       // 32 keys -> 16 packed columns per tcgen05.st
          uint32_t pk[16];
          exp_keys(c * 32, 32, pk);
          tmem_st_32x16(t_p + c * 16, pk);
          if (more) tmem_ld_32x32(t_s + c * 32, v + c * 32);
        }
        const float2 t = __fadd2_rn(__fadd2_rn(ls[0], ls[1]), __fadd2_rn(ls[2], ls[3]));
        l_run += t.x + t.y;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_full);
        if (more) {
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(s_empty);                             // tile j+1's scores are in registers: Q K^T of tile j+2 may start
        }
      }
    } else
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kbase = j * TA_BK;
      const bool tail = kbase + TA_BK > p.Sk;
      const bool diag = p.causal && (kbase + TA_BK - 1 > q0);   // some key of this tile lies after some query of the tile
      uint32_t v[128];
      tmem_ld_32x32(t_s, v);
      tmem_ld_32x32(t_s + 32, v + 32);
      tmem_ld_32x32(t_s + 64, v + 64);
      tmem_ld_32x32(t_s + 96, v + 96);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_empty);                                 // scores are in registers: QK^T of the next tile may start
      if (tail || diag) {
        const int lim = diag ? min(p.Sk, q0 + row + 1) : p.Sk;   // first masked key index for this query row
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (kbase + i >= lim) v[i] = __float_as_uint(-INFINITY);
      }
      // Row maximum: computed for the first tile only.  Later tiles run their exponentials against the running reference
      // maximum m_ref straight away and look at the tile's row sum afterwards: a sum above 2^15 (some score more than
      // ~2^8 above the reference; +inf on overflow) sends the row through the slow path -- true maximum, O and l brought
      // to the new reference, exponentials redone (P has not been handed to the MMA warp yet).  That is the same "lazy
      // rescale" rule as before, minus 128 FMNMX and their dependent chain on every tile's critical path.
      auto row_max = [&]() {
        float mxp[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) mxp[c] = __uint_as_float(v[c]);
#pragma unroll
        for (int i = 8; i < 128; i += 8) {
#pragma unroll
          for (int c = 0; c < 8; ++c) mxp[c] = fmaxf(mxp[c], __uint_as_float(v[i + c]));
        }
        return fmaxf(fmaxf(fmaxf(mxp[0], mxp[1]), fmaxf(mxp[2], mxp[3])), fmaxf(fmaxf(mxp[4], mxp[5]), fmaxf(mxp[6], mxp[7])));
      };
      auto exp_tile = [&](float msc) {                      // P = 2^(s * scale_log2 - msc) -> TMEM; returns the row sum
        // packed fp32x2 arithmetic (sm_100 FFMA2 / FADD2): one instruction scales two scores, one accumulates two
        // exponentials -- 2.5 instructions per exponential instead of 3.5 around the 8-cycle MUFU cadence
        const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(-msc, -msc);
        float2 ls[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int c = 0; c < 2; ++c) {                       // 64 keys -> 32 packed columns per tcgen05.st
          uint32_t pk[32];
#pragma unroll
          for (int i = 0; i < 64; i += 2) {
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(v[c * 64 + i]), __uint_as_float(v[c * 64 + i + 1])), sc2, nm2);
            float2 e;
            if ((((i >> 1) * NP8) & 7) < NP8) {
              e = poly_exp2_pair(x);
            } else {
              e.x = fast_exp2(x.x);
              e.y = fast_exp2(x.y);
            }
            ls[(i >> 1) & 3] = __fadd2_rn(ls[(i >> 1) & 3], e);
            pk[i >> 1] = pack_bf16x2(e.x, e.y);
          }
          if (DEFER && c == 0) {
            mbar_wait(p_empty, (j & 1) ^ 1);                // P V of the previous tile has consumed P (and updated O)
            tc_fence_after();
          }
          tmem_st_32x32(t_p + c * 32, pk);
        }
        const float2 t = __fadd2_rn(__fadd2_rn(ls[0], ls[1]), __fadd2_rn(ls[2], ls[3]));
        return t.x + t.y;
      };
      if (j == 0) m_ref = row_max();
      if (!DEFER) {
        mbar_wait(p_empty, (j & 1) ^ 1);                    // P V of the previous tile has consumed P (and updated O)
        tc_fence_after();
      }
      float tile_sum = exp_tile(m_ref * p.scale_log2);
      if (__any_sync(0xffffffffu, !(tile_sum <= 32768.0f))) {
        // rare, and taken by the whole warp (tcgen05.ld/st are warp-collective): bring O (TMEM) and l to the new reference
        // maximum -- rows that did not need it keep theirs (f == 1); o_full(j-1) completed together with p_empty above
        const float mx = fmaxf(m_ref, row_max());
        const float f = fast_exp2((m_ref - mx) * p.scale_log2);
        tmem_st_wait();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t o[32];
          tmem_ld_32x32(t_o + c * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
          tmem_st_32x32(t_o + c * 32, o);
        }
        l_run *= f;
        m_ref = mx;
        tile_sum = exp_tile(m_ref * p.scale_log2);
      }
      tmem_st_wait();
      l_run += tile_sum;
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // epilogue: O / l  ->  bf16 tile through sQ (Q is dead: every QK^T has completed) -> TMA store clips rows >= Sq
    mbar_wait(o_full, (n_kv - 1) & 1);
    tc_fence_after();
    const float inv = 1.f / l_run;
    const uint32_t sO_row = smem_u32(sQ) + row * 128;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld_32x32(t_o + c * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const float* f = reinterpret_cast<const float*>(o) + ch * 8;
        st_shared_v4(sO_row + (((c * 4 + ch) ^ sw) << 4), pack_bf16x2(f[0] * inv, f[1] * inv), pack_bf16x2(f[2] * inv, f[3] * inv),
                     pack_bf16x2(f[4] * inv, f[5] * inv), pack_bf16x2(f[6] * inv, f[7] * inv));
      }
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (threadIdx.x == 0) {
      tma_store_3d(&tmap_o, sQ, h * TA_HD, q0, b);
      tma_store_commit();
      tma_store_wait_all<0>();
    }
    if (p.lse != nullptr && q0 + row < p.Sq)
      p.lse[((int64_t)b * p.H + h) * p.Sq + q0 + row] = m_ref * p.scale + __logf(l_run);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TA_TMEM_COLS);
  }
}

// [B, S, cols] view of a [B*S, ld] matrix as a 3-D tensor map; box = 64 columns x 128 rows x 1 batch, 128B swizzle
static int make_tmap_bsc(CUtensorMap* out, const void* base, int64_t ld, int B, int S, int cols, uint32_t box_rows = 128) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) { dwb_set_error("cuTensorMapEncodeTiled entry point unavailable"); return DWB_ERR_CUDA; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((ld * 2) & 15) != 0) {
    dwb_set_error("attention operand needs a 16 B aligned base and row pitch (base=%p pitch=%lld B)", base, (long long)ld * 2);
    return DWB_ERR_INVALID;
  }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)S, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)S * (cuuint64_t)ld * 2};
  cuuint32_t box[3] = {64, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { dwb_set_error("cuTensorMapEncodeTiled(3d) failed with %d", (int)r); return DWB_ERR_CUDA; }
  return DWB_OK;
}

}  // namespace dwb

using namespace dwb;

extern "C" int dwb_attention_fwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                                    int64_t ldo, float* lse, int B, int H, int Sq, int Sk, int head_dim, int causal, float scale,
                                    void* stream) {
  DWB_CHECK_ARG(head_dim == TA_HD, "dwb_attention_fwd_tc: head_dim %d unsupported (Whisper uses 64)", head_dim);
  DWB_CHECK_ARG(q && k && v && o, "dwb_attention_fwd_tc: null operand");
  DWB_CHECK_ARG(B > 0 && H > 0 && Sq > 0 && Sk > 0, "dwb_attention_fwd_tc: bad shape");
  CUtensorMap tq, tk, tv, to;
  int rc;
  if ((rc = make_tmap_bsc(&tq, q, ldq, B, Sq, H * TA_HD))) return rc;
  if ((rc = make_tmap_bsc(&tk, k, ldk, B, Sk, H * TA_HD))) return rc;
  if ((rc = make_tmap_bsc(&tv, v, ldv, B, Sk, H * TA_HD))) return rc;
  if ((rc = make_tmap_bsc(&to, o, ldo, B, Sq, H * TA_HD))) return rc;
  // DWB_ATTN_POLY = eighths of the exponentials emulated on the FMA pipe (0..4) + 10 * mode (0 original loop, 1 deferred P wait,
  // 2 software-pipelined loop): A/B switch for the microbenchmarks; the default is the measured best (profiles/)
  static const int variant = [] { const char* e = getenv("DWB_ATTN_POLY"); return e ? atoi(e) : TA_DEFAULT_VARIANT; }();
  const int np8 = variant % 10 > 4 ? 4 : variant % 10;
  const int mode = variant / 10 > 2 ? 2 : variant / 10;
  typedef void (*kern_t)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const TcAttnParams);
  static const kern_t table[3][5] = {
      {attn_fwd_tc_kernel<0, 0>, attn_fwd_tc_kernel<1, 0>, attn_fwd_tc_kernel<2, 0>, attn_fwd_tc_kernel<3, 0>, attn_fwd_tc_kernel<4, 0>},
      {attn_fwd_tc_kernel<0, 1>, attn_fwd_tc_kernel<1, 1>, attn_fwd_tc_kernel<2, 1>, attn_fwd_tc_kernel<3, 1>, attn_fwd_tc_kernel<4, 1>},
      {attn_fwd_tc_kernel<0, 2>, attn_fwd_tc_kernel<1, 2>, attn_fwd_tc_kernel<2, 2>, attn_fwd_tc_kernel<3, 2>, attn_fwd_tc_kernel<4, 2>}};
  static bool attr = false;
  if (!attr) {
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 5; ++b) DWB_CUDA_OK(cudaFuncSetAttribute(table[a][b], cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM));
    attr = true;
  }
  TcAttnParams p;
  p.H = H; p.Sq = Sq; p.Sk = Sk; p.causal = causal;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse = lse;
  p.reverse = dwb_row_walk_reverse();
  dim3 grid(ceil_div(Sq, TA_BQ), B * H);
  // (Two variants with two softmax threads per query row -- 16 softmax warps per SM -- were built and measured in round 1:
  //  with a per-tile pair barrier 0.70 ms, with each thread reducing the full-row maximum itself (no exchange) 0.62 ms,
  //  against 0.56 ms for this kernel: the XU pipe is busy 63 % of the time here, but the extra TMEM traffic and the
  //  tighter register budget (104/thread) cost more than the additional warps recover.  A third variant with 64-key tiles
  //  and double-buffered S / P in TMEM (QK^T two tiles ahead, no wait on P V) measured 0.60 ms: per-tile instruction and
  //  barrier overhead, not the dependency chain, is what the power-capped SM pays for.  A fourth, FA4-shaped kernel -- one
  //  CTA per SM owning two query tiles (512 TMEM columns), K/V loaded once for both, optional named-barrier ping-pong of the
  //  two softmax warpgroups -- measured 0.63 ms with and without the ping-pong: one MMA-issuing thread serving two tiles in
  //  program order and a single softmax warp per SM sub-partition on the XU pipe at a time are both worse than two
  //  independent CTAs.  See profiles/README.md.)
  table[mode][np8]<<<grid, TA_THREADS, TA_SMEM, (cudaStream_t)stream>>>(tq, tk, tv, to, p);
  DWB_LAUNCH_OK();
  return DWB_OK;
}
