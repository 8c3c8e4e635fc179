// Non-causal attention forward on tcgen05 / TMEM (head_dim 64): the Whisper encoder self-attention, S = 1500.
//   per CTA: one (batch, head, 128-query tile); loops over 128-key tiles
//   S = Q K^T      tcgen05.mma 128x128x64  (Q, K tiles: TMA, 128B swizzle, K-major)        -> TMEM cols [0,128)
//   softmax        one thread per query row (tcgen05.ld 32x32b: lane == row, no shuffles), online max/sum in fp32,
//                  P (bf16) written to a swizzled smem tile
//   O_j = P V      tcgen05.mma 128x64x128  (V tile is the MN-major B operand: no transpose)  -> TMEM cols [128,192)
//   O  += O_j      folded into registers with the running-max correction, normalised at the end, TMA store
// Two CTAs are resident per SM (112 KB smem, 256 TMEM columns each), so one CTA's exp2-bound softmax overlaps the
// other's MMAs; inside a CTA, QK^T of tile j+1 is issued as soon as tile j's scores have left TMEM.
//
// Replaces F.scaled_dot_product_attention (HF:integrations/sdpa_attention.py:40-104) for
// HF:models/whisper/modeling_whisper.py:342-352 in the encoder (is_causal False), same contract as dwb_attention_fwd.
#include "common.cuh"

namespace dwb {

constexpr int TA_BQ = 128, TA_BK = 128, TA_HD = 64;
constexpr int TA_THREADS = 192;
constexpr int TA_TILE_BYTES = 128 * 128;            // 128 rows x 128 B
constexpr int TA_TILES_BYTES = TA_TILE_BYTES /*Q*/ + 2 * 2 * TA_TILE_BYTES /*K,V x 2 stages*/ + 2 * TA_TILE_BYTES /*P*/;
constexpr int TA_BAR_BYTES = 96;
// two CTAs per SM: 2 * (dyn + 1 KB system reserve) <= 228 KB  ->  dyn <= 115712; the slack absorbs the 1 KB round-up
constexpr int TA_SMEM = 115712;
static_assert(TA_TILES_BYTES + TA_BAR_BYTES + 896 <= TA_SMEM, "smem budget");
constexpr int TA_TMEM_COLS = 256;

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

struct TcAttnParams {
  int H, Sq, Sk;
  float scale_log2;     // scale * log2(e)
  float scale;
  float* lse;           // [B, H, Sq] or null
};

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
  uint8_t* sP = sKV + 4 * TA_TILE_BYTES;              // two 128x64 halves
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * TA_TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;      // [2]
  uint64_t* kv_empty = bars + 3;     // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_empty = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* p_empty = bars + 8;
  uint64_t* o_full = bars + 9;
  uint64_t* o_empty = bars + 10;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * TA_BQ;
  const int b = blockIdx.y / p.H, h = blockIdx.y % p.H;
  const int n_kv = ceil_div(p.Sk, TA_BK);

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v); tma_prefetch_desc(&tmap_o);
    mbar_init(q_full, 1);
    mbar_init(&kv_full[0], 1); mbar_init(&kv_full[1], 1);
    mbar_init(&kv_empty[0], 1); mbar_init(&kv_empty[1], 1);
    mbar_init(s_full, 1); mbar_init(s_empty, 128);
    mbar_init(p_full, 128); mbar_init(p_empty, 1);
    mbar_init(o_full, 1); mbar_init(o_empty, 128);
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
  const uint32_t tmem_s = tmem_base;            // 128 columns of scores
  const uint32_t tmem_o = tmem_base + 128;      // 64 columns: P V of the current tile

  if (warp == 4) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      mbar_expect_tx(q_full, TA_TILE_BYTES);
      tma_load_3d(&tmap_q, q_full, sQ, h * TA_HD, q0, b);
    }
    for (int j = 0; j < n_kv; ++j) {
      const int st = j & 1;
      mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
      if (lane == 0) {
        mbar_expect_tx(&kv_full[st], 2 * TA_TILE_BYTES);
        tma_load_3d(&tmap_k, &kv_full[st], sKV + st * 2 * TA_TILE_BYTES, h * TA_HD, j * TA_BK, b);
        tma_load_3d(&tmap_v, &kv_full[st], sKV + st * 2 * TA_TILE_BYTES + TA_TILE_BYTES, h * TA_HD, j * TA_BK, b);
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    // ===================================== MMA issuer =======================================
    constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);
    mbar_wait(q_full, 0);
    const uint64_t dq = umma_desc_sw128(smem_u32(sQ), 16, 1024);
    for (int j = 0; j < n_kv; ++j) {
      const int st = j & 1;
      mbar_wait(&kv_full[st], (j >> 1) & 1);
      mbar_wait(s_empty, (j & 1) ^ 1);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t dk = umma_desc_sw128(smem_u32(sKV + st * 2 * TA_TILE_BYTES), 16, 1024);
#pragma unroll
        for (int k = 0; k < TA_HD / 16; ++k) tc_mma_ss(tmem_s, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_qk, k > 0 ? 1u : 0u);
        tc_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, j & 1);
      mbar_wait(o_empty, (j & 1) ^ 1);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t dv = umma_desc_sw128(smem_u32(sKV + st * 2 * TA_TILE_BYTES + TA_TILE_BYTES), TA_TILE_BYTES, 1024);
#pragma unroll
        for (int k = 0; k < TA_BK / 16; ++k) {
          const uint64_t dp = umma_desc_sw128(smem_u32(sP + (k >> 2) * TA_TILE_BYTES), 16, 1024) + (uint64_t)(2 * (k & 3));
          tc_mma_ss(tmem_o, dp, dv + (uint64_t)(k * 128), idesc_pv, k > 0 ? 1u : 0u);
        }
        tc_commit(o_full);
        tc_commit(&kv_empty[st]);
        tc_commit(p_empty);
      }
      __syncwarp();
    }
  } else {
    // ===================================== softmax / output =================================
    const int row = warp * 32 + lane;                       // query row of this thread == TMEM lane
    const uint32_t t_s = tmem_s + ((uint32_t)(warp * 32) << 16);
    const uint32_t t_o = tmem_o + ((uint32_t)(warp * 32) << 16);
    const uint32_t sP_row = smem_u32(sP) + row * 128;
    const int sw = row & 7;
    float m_run = -INFINITY, l_run = 0.f, m_prev = -INFINITY, m_ref = -INFINITY;
    float o_acc[TA_HD];
#pragma unroll
    for (int i = 0; i < TA_HD; ++i) o_acc[i] = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kbase = j * TA_BK;
      const bool tail = kbase + TA_BK > p.Sk;
      // pass 1: row max
      float mx = m_run;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(t_s + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float s = __uint_as_float(v[i]);
          if (tail && kbase + c * 32 + i >= p.Sk) s = -INFINITY;
          mx = fmaxf(mx, s);
        }
      }
      const float m_new = mx;                               // finite: every tile holds at least one valid key
      const float msc = m_new * p.scale_log2;
      const float corr = (m_run == -INFINITY) ? 0.f : exp2f(m_run * p.scale_log2 - msc);
      l_run *= corr;
      mbar_wait(p_empty, (j & 1) ^ 1);                      // P V of the previous tile has consumed sP
      // pass 2: p = exp2(s * scale_log2 - m * scale_log2), row sum, bf16 pack into the swizzled P tile
      float lsum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(t_s + c * 32, v);
        tmem_ld_wait();
        if (c == 3) {                                       // scores fully read: QK^T of the next tile may overwrite
          tc_fence_before();
          mbar_arrive(s_empty);
        }
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float e0 = exp2f(__uint_as_float(v[i]) * p.scale_log2 - msc);
          float e1 = exp2f(__uint_as_float(v[i + 1]) * p.scale_log2 - msc);
          if (tail) {
            if (kbase + c * 32 + i >= p.Sk) e0 = 0.f;
            if (kbase + c * 32 + i + 1 >= p.Sk) e1 = 0.f;
          }
          lsum += e0 + e1;
          pk[i >> 1] = pack_bf16x2(e0, e1);
        }
        const uint32_t half = sP_row + (c >> 1) * TA_TILE_BYTES;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int chunk = (c & 1) * 4 + q4;               // 16 B chunk (8 keys) inside the 128 B half-row
          st_shared_v4(half + ((chunk ^ sw) << 4), pk[q4 * 4], pk[q4 * 4 + 1], pk[q4 * 4 + 2], pk[q4 * 4 + 3]);
        }
      }
      l_run += lsum;
      fence_proxy_async_smem();
      mbar_arrive(p_full);
      // fold the previous tile's P V (computed against m_prev) into the register accumulator
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
        const float f = (m_ref == -INFINITY) ? 0.f : exp2f((m_ref - m_prev) * p.scale_log2);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(t_o + c * 32, v);
          tmem_ld_wait();
          if (c == 1) {
            tc_fence_before();
            mbar_arrive(o_empty);
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = fmaf(o_acc[c * 32 + i], f, __uint_as_float(v[i]));
        }
        m_ref = m_prev;
      }
      m_prev = m_new;
      m_run = m_new;
    }
    // last tile's P V
    {
      mbar_wait(o_full, (n_kv - 1) & 1);
      tc_fence_after();
      const float f = (m_ref == -INFINITY) ? 0.f : exp2f((m_ref - m_prev) * p.scale_log2);
      const float inv = 1.f / l_run;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(t_o + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = fmaf(o_acc[c * 32 + i], f, __uint_as_float(v[i])) * inv;
      }
    }
    // output tile through sQ (Q is dead: every QK^T has completed) -> TMA store clips rows >= Sq
    const uint32_t sO_row = smem_u32(sQ) + row * 128;
#pragma unroll
    for (int ch = 0; ch < 8; ++ch)
      st_shared_v4(sO_row + ((ch ^ sw) << 4), pack_bf16x2(o_acc[ch * 8], o_acc[ch * 8 + 1]), pack_bf16x2(o_acc[ch * 8 + 2], o_acc[ch * 8 + 3]),
                   pack_bf16x2(o_acc[ch * 8 + 4], o_acc[ch * 8 + 5]), pack_bf16x2(o_acc[ch * 8 + 6], o_acc[ch * 8 + 7]));
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (threadIdx.x == 0) {
      tma_store_3d(&tmap_o, sQ, h * TA_HD, q0, b);
      tma_store_commit();
      tma_store_wait_all<0>();
    }
    if (p.lse != nullptr && q0 + row < p.Sq)
      p.lse[((int64_t)b * p.H + h) * p.Sq + q0 + row] = m_run * p.scale + __logf(l_run);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TA_TMEM_COLS);
  }
}

// [B, S, cols] view of a [B*S, ld] matrix as a 3-D tensor map; box = 64 columns x 128 rows x 1 batch, 128B swizzle
static int make_tmap_bsc(CUtensorMap* out, const void* base, int64_t ld, int B, int S, int cols) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) { dwb_set_error("cuTensorMapEncodeTiled entry point unavailable"); return DWB_ERR_CUDA; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((ld * 2) & 15) != 0) {
    dwb_set_error("attention operand needs a 16 B aligned base and row pitch (base=%p pitch=%lld B)", base, (long long)ld * 2);
    return DWB_ERR_INVALID;
  }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)S, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)S * (cuuint64_t)ld * 2};
  cuuint32_t box[3] = {64, 128, 1};
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
  DWB_CHECK_ARG(!causal, "dwb_attention_fwd_tc: causal attention uses dwb_attention_fwd");
  DWB_CHECK_ARG(q && k && v && o, "dwb_attention_fwd_tc: null operand");
  DWB_CHECK_ARG(B > 0 && H > 0 && Sq > 0 && Sk > 0, "dwb_attention_fwd_tc: bad shape");
  CUtensorMap tq, tk, tv, to;
  int rc;
  if ((rc = make_tmap_bsc(&tq, q, ldq, B, Sq, H * TA_HD))) return rc;
  if ((rc = make_tmap_bsc(&tk, k, ldk, B, Sk, H * TA_HD))) return rc;
  if ((rc = make_tmap_bsc(&tv, v, ldv, B, Sk, H * TA_HD))) return rc;
  if ((rc = make_tmap_bsc(&to, o, ldo, B, Sq, H * TA_HD))) return rc;
  static bool attr = false;
  if (!attr) {
    DWB_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM));
    attr = true;
  }
  TcAttnParams p;
  p.H = H; p.Sq = Sq; p.Sk = Sk;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse = lse;
  dim3 grid(ceil_div(Sq, TA_BQ), B * H);
  attn_fwd_tc_kernel<<<grid, TA_THREADS, TA_SMEM, (cudaStream_t)stream>>>(tq, tk, tv, to, p);
  DWB_LAUNCH_OK();
  return DWB_OK;
}
