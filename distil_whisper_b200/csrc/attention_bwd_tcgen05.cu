// Attention backward on tcgen05 / TMEM (head_dim 64): gradients of softmax(scale Q K^T [causal]) V.
//   per CTA: one (batch, head, 128-key tile j); loops over the 128-query tiles i that see it.  All five contractions of
//   flash-attention backward run on the tensor core, transposed so that TMEM lanes are KEY rows:
//     S^T  = K_j Q_i^T          (SS, 128x128x64)   -> TMEM [  0,128)
//     dP^T = V_j dO_i^T         (SS, 128x128x64)   -> TMEM [128,256)
//     P^T  = exp2(S^T c - lse_q),  dS^T = P^T (dP^T - delta_q) scale      (two threads per key row, 64 query columns each, fp32)
//     dV_j += P^T  dO_i         (TS: P^T from TMEM [256,320) as bf16 pairs; dO_i is the MN-major B operand)   -> TMEM [320,384)
//     dK_j += dS^T Q_i          (SS: dS^T tile in smem, K-major A; Q_i MN-major B)                            -> TMEM [384,448)
//     dQ_i  = dS   K_j          (SS: the SAME smem tile read as an MN-major A operand; K_j MN-major B)        -> TMEM [448,512)
//   dQ_i partial tiles are reduced across key tiles with TMA reduce-add (fp32) into dq_acc; dK_j / dV_j leave through TMA.
//   No operand is transposed in memory: the UMMA major bits select the orientation.
//
// Same contract as dwb_attention_bwd (autograd of HF:integrations/sdpa_attention.py:40-104 as reached from
// HF:models/whisper/modeling_whisper.py:342-352 under ref:training/run_distillation.py:1609).
#include "common.cuh"

namespace dwb {

constexpr int AB_T = 128;                   // tile edge (queries and keys)
constexpr int AB_HD = 64;
constexpr int AB_TILE = 128 * 128;          // bytes of a [128 x 64] bf16 tile
constexpr int AB_THREADS = 384;             // warps 0-7 compute (two threads per key row: 64 query columns each), warp 8 TMA, warp 9 MMA
constexpr int AB_OFF_K = 0, AB_OFF_V = AB_TILE, AB_OFF_Q = 2 * AB_TILE /*2 stages*/, AB_OFF_DO = 4 * AB_TILE /*2 stages*/,
              AB_OFF_DS = 6 * AB_TILE /*2 halves*/, AB_OFF_DQ = 8 * AB_TILE /*fp32 staging, 2 panels of 16 KB*/,
              AB_OFF_END = 10 * AB_TILE;
constexpr int AB_SMEM = 1024 + AB_OFF_END + 4 * 128 * 4 /*lse, delta x 2 stages*/ + 256;
constexpr uint32_t AB_TM_S = 0, AB_TM_DP = 128, AB_TM_P = 256, AB_TM_DV = 320, AB_TM_DK = 384, AB_TM_DQ = 448;

__device__ __forceinline__ void tma_store_3d_b(const CUtensorMap* m, const void* smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* m, const void* smem, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

struct TcAttnBwdParams {
  int H, Sq, Sk, causal;
  float scale, scale_log2;
  const float* lse;      // [B, H, Sq]
  const float* delta;    // [B, H, Sq]
};

__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                   const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                   const __grid_constant__ CUtensorMap tmap_dq /*fp32*/, const __grid_constant__ CUtensorMap tmap_dk,
                   const __grid_constant__ CUtensorMap tmap_dv, const TcAttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem + AB_OFF_K;
  uint8_t* sV = smem + AB_OFF_V;
  uint8_t* sQ = smem + AB_OFF_Q;
  uint8_t* sDO = smem + AB_OFF_DO;
  uint8_t* sDS = smem + AB_OFF_DS;
  uint8_t* sDQ = smem + AB_OFF_DQ;
  float* sLse = reinterpret_cast<float*>(smem + AB_OFF_END);      // [2][128]  lse * log2(e)
  float* sDel = sLse + 256;                                       // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDel + 256);
  uint64_t* kv_full = bars;
  uint64_t* q_full = bars + 1;       // [2]
  uint64_t* q_empty = bars + 3;      // [2]
  uint64_t* sdp_full = bars + 5;
  uint64_t* sdp_empty = bars + 6;
  uint64_t* pds_full = bars + 7;
  uint64_t* pds_empty = bars + 8;
  uint64_t* dq_full = bars + 9;
  uint64_t* dq_empty = bars + 10;
  uint64_t* acc_full = bars + 11;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * AB_T;
  const int b = blockIdx.y / p.H, h = blockIdx.y % p.H;
  const int n_q = ceil_div(p.Sq, AB_T);
  const int i0 = p.causal ? min(k0 / AB_T, n_q) : 0;     // first query tile that sees this key tile
  const int n_it = n_q - i0;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v); tma_prefetch_desc(&tmap_do);
    tma_prefetch_desc(&tmap_dq); tma_prefetch_desc(&tmap_dk); tma_prefetch_desc(&tmap_dv);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    mbar_init(sdp_full, 1); mbar_init(sdp_empty, 256);
    mbar_init(pds_full, 256); mbar_init(pds_empty, 1);
    mbar_init(dq_full, 1); mbar_init(dq_empty, 256);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 8) {
    // ===================================== TMA producer (warp-uniform loop, elected lane issues) ==========
    if (n_it > 0) {
      const bool leader = elect_one();
      if (leader) {
        mbar_expect_tx(kv_full, 2 * AB_TILE);
        tma_load_3d(&tmap_k, kv_full, sK, h * AB_HD, k0, b);
        tma_load_3d(&tmap_v, kv_full, sV, h * AB_HD, k0, b);
      }
      for (int it = 0; it < n_it; ++it) {
        const int st = it & 1;
        mbar_wait(&q_empty[st], ((it >> 1) & 1) ^ 1);
        if (leader) {
          mbar_expect_tx(&q_full[st], 2 * AB_TILE);
          tma_load_3d(&tmap_q, &q_full[st], sQ + st * AB_TILE, h * AB_HD, (i0 + it) * AB_T, b);
          tma_load_3d(&tmap_do, &q_full[st], sDO + st * AB_TILE, h * AB_HD, (i0 + it) * AB_T, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 9) {
    // ===================================== MMA issuer (warp-uniform loop, elected lane issues) ============
    if (n_it > 0) {
      const bool leader = elect_one();
      constexpr uint32_t id_nn = umma_idesc_bf16(128, 128, 0, 0);     // S^T, dP^T
      constexpr uint32_t id_kb = umma_idesc_bf16(128, 64, 0, 1);      // dV (TS), dK: K-major A, MN-major B
      constexpr uint32_t id_mm = umma_idesc_bf16(128, 64, 1, 1);      // dQ: MN-major A and B
      mbar_wait(kv_full, 0);
      const uint64_t dK_k = umma_desc_sw128(smem_u32(sK), 16, 1024);            // K_j as K-major A
      const uint64_t dV_k = umma_desc_sw128(smem_u32(sV), 16, 1024);            // V_j as K-major A
      const uint64_t dK_mn = umma_desc_sw128(smem_u32(sK), AB_TILE, 1024);      // K_j as MN-major B
      const uint64_t dDS_k = umma_desc_sw128(smem_u32(sDS), 16, 1024);          // dS^T as K-major A (two halves)
      const uint64_t dDS_mn = umma_desc_sw128(smem_u32(sDS), AB_TILE, 1024);    // dS as MN-major A (query groups 16 KB apart)
      auto issue_sdp = [&](int it) {
        const int st = it & 1;
        mbar_wait(&q_full[st], (it >> 1) & 1);
        mbar_wait(sdp_empty, (it & 1) ^ 1);
        tc_fence_after();
        if (leader) {
          const uint64_t dQ_k = umma_desc_sw128(smem_u32(sQ + st * AB_TILE), 16, 1024);
          const uint64_t dDO_k = umma_desc_sw128(smem_u32(sDO + st * AB_TILE), 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_ss(tmem_base + AB_TM_S, dK_k + (uint64_t)(2 * k), dQ_k + (uint64_t)(2 * k), id_nn, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_ss(tmem_base + AB_TM_DP, dV_k + (uint64_t)(2 * k), dDO_k + (uint64_t)(2 * k), id_nn, k > 0 ? 1u : 0u);
          tc_commit(sdp_full);
        }
        __syncwarp();
      };
      issue_sdp(0);
      for (int it = 0; it < n_it; ++it) {
        const int st = it & 1;
        if (it + 1 < n_it) issue_sdp(it + 1);
        mbar_wait(pds_full, it & 1);
        tc_fence_after();
        const uint64_t dQ_mn = umma_desc_sw128(smem_u32(sQ + st * AB_TILE), AB_TILE, 1024);
        const uint64_t dDO_mn = umma_desc_sw128(smem_u32(sDO + st * AB_TILE), AB_TILE, 1024);
        if (leader) {
#pragma unroll
          for (int k = 0; k < 8; ++k)     // dV += P^T dO_i      (K = 128 queries, 16 per step = 8 TMEM columns / 2048 B of dO)
            tc_mma_ts(tmem_base + AB_TM_DV, tmem_base + AB_TM_P + 8 * k, dDO_mn + (uint64_t)(k * 128), id_kb, (it > 0 || k > 0) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 8; ++k)     // dK += dS^T Q_i
            tc_mma_ss(tmem_base + AB_TM_DK, dDS_k + (uint64_t)((k >> 2) * (AB_TILE >> 4) + 2 * (k & 3)), dQ_mn + (uint64_t)(k * 128), id_kb,
                      (it > 0 || k > 0) ? 1u : 0u);
        }
        __syncwarp();
        mbar_wait(dq_empty, (it & 1) ^ 1);   // the compute threads drain tile it-1's dQ while dV / dK above run
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int k = 0; k < 8; ++k)     // dQ_i = dS K_j       (K = 128 keys)
            tc_mma_ss(tmem_base + AB_TM_DQ, dDS_mn + (uint64_t)(k * 128), dK_mn + (uint64_t)(k * 128), id_mm, k > 0 ? 1u : 0u);
          tc_commit(dq_full);
          tc_commit(&q_empty[st]);
          tc_commit(pds_empty);
        }
        __syncwarp();
      }
      if (leader) tc_commit(acc_full);
    }
  } else if (warp < 8) {
    // ===================================== compute: two threads per key row =================
    const int wq = warp & 3, hsel = warp >> 2;       // TMEM lane quadrant, and which 64 query columns this thread owns
    const int row = wq * 32 + lane;                   // key row inside the tile == TMEM lane
    const int key = k0 + row;
    const bool key_ok = key < p.Sk;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const uint32_t sDS_row = smem_u32(sDS) + row * 128;
    const uint32_t sDQ_row = smem_u32(sDQ) + row * 128;
    const int sw = row & 7;
    const float* LSE = p.lse + ((int64_t)b * p.H + h) * p.Sq;
    const float* DEL = p.delta + ((int64_t)b * p.H + h) * p.Sq;

    // dQ_i partial tile: TMEM -> fp32 smem panels -> TMA reduce-add into dq_acc (TMEM lanes are QUERY rows for this accumulator)
    auto drain_dq = [&](int t) {
      mbar_wait(dq_full, t & 1);
      tc_fence_after();
      if (threadIdx.x == 0) tma_store_wait_read<0>();           // previous reduce-add has left the staging panels
      named_bar_sync(2, 256);
      {
        const int pnl = hsel;                                   // this warpgroup's 32 fp32 columns
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + AB_TM_DQ + lane_off + pnl * 32, v);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(dq_empty);
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
          st_shared_v4(sDQ_row + pnl * AB_TILE + ((ch ^ sw) << 4), v[ch * 4], v[ch * 4 + 1], v[ch * 4 + 2], v[ch * 4 + 3]);
      }
      fence_proxy_async_smem();
      named_bar_sync(3, 256);
      if (threadIdx.x == 0) {
        const int tq0 = (i0 + t) * AB_T;
        tma_reduce_add_3d(&tmap_dq, sDQ, h * AB_HD, tq0, b);
        tma_reduce_add_3d(&tmap_dq, sDQ + AB_TILE, h * AB_HD + 32, tq0, b);
        tma_store_commit();
      }
    };

    float pre_lse = INFINITY, pre_del = 0.f;
    if (n_it > 0) {
      const int qi = i0 * AB_T + row;
      if (qi < p.Sq) { pre_lse = __ldg(LSE + qi) * 1.4426950408889634f; pre_del = __ldg(DEL + qi); }
    }
    for (int it = 0; it < n_it; ++it) {
      const int st = it & 1;
      const int q0 = (i0 + it) * AB_T;
      // per-query constants of this tile: lse (in log2 units) and delta; +inf lse zeroes padded query columns.
      // They were fetched one iteration ahead (pre_lse / pre_del), so the global-load latency is off the critical path.
      if (hsel == 0) {
        sLse[st * 128 + row] = pre_lse;
        sDel[st * 128 + row] = pre_del;
      }
      {
        const int qn = q0 + AB_T + row;                         // next tile's query handled by this thread
        const bool ok = (it + 1 < n_it) && qn < p.Sq;
        pre_lse = ok ? __ldg(LSE + qn) * 1.4426950408889634f : INFINITY;
        pre_del = ok ? __ldg(DEL + qn) : 0.f;
      }
      named_bar_sync(1, 256);
      mbar_wait(sdp_full, it & 1);
      tc_fence_after();
      const bool diag = p.causal && (q0 < k0 + AB_T);          // tile touches the causal boundary
#pragma unroll 1
      for (int qt = 2 * hsel; qt < 2 * hsel + 2; ++qt) {        // this thread's 64 query columns, 32 at a time
        uint32_t s[32], dp[32];
        tmem_ld_32x32(tmem_base + AB_TM_S + lane_off + qt * 32, s);
        tmem_ld_32x32(tmem_base + AB_TM_DP + lane_off + qt * 32, dp);
        tmem_ld_wait();
        if (qt == 2 * hsel + 1) {                               // this thread's scores are read: S^T / dP^T of the next tile may land
          tc_fence_before();
          mbar_arrive(sdp_empty);
        }
        uint32_t pk[16], dk[16];
#pragma unroll
        for (int q = 0; q < 32; q += 2) {
          float pv[2], dv[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int qc = qt * 32 + q + e;
            const float l2 = sLse[st * 128 + qc];
            float pe = fast_exp2(fmaf(__uint_as_float(s[q + e]), p.scale_log2, -l2));
            if (!key_ok || (diag && key > q0 + qc)) pe = 0.f;
            pv[e] = pe;
            dv[e] = pe * (__uint_as_float(dp[q + e]) - sDel[st * 128 + qc]) * p.scale;
          }
          pk[q >> 1] = pack_bf16x2(pv[0], pv[1]);
          dk[q >> 1] = pack_bf16x2(dv[0], dv[1]);
        }
        if (qt == 2 * hsel) {                                   // P^T (TMEM) and dS^T (smem) of the previous tile consumed?
          mbar_wait(pds_empty, (it & 1) ^ 1);                   // (waited for only now: the math above overlapped the MMAs)
          tc_fence_after();
        }
        tmem_st_32x16(tmem_base + AB_TM_P + lane_off + qt * 16, pk);
        const uint32_t half_base = sDS_row + (qt >> 1) * AB_TILE;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)                         // 32 queries = 64 B = 4 chunks of this key's dS^T row
          st_shared_v4(half_base + ((((qt & 1) * 4 + ch) ^ sw) << 4), dk[ch * 4], dk[ch * 4 + 1], dk[ch * 4 + 2], dk[ch * 4 + 3]);
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(pds_full);
      if (it > 0) drain_dq(it - 1);        // previous tile's dQ: its MMAs finished long ago; runs under this tile's dV / dK MMAs
    }
    if (n_it > 0) drain_dq(n_it - 1);
    // dV_j, dK_j: TMEM -> bf16 tiles -> TMA store (clips key rows >= Sk)
    if (n_it > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
    }
    if (threadIdx.x == 0) tma_store_wait_read<0>();
    named_bar_sync(2, 256);
    {
      const int which = hsel;                                   // warpgroup 0 drains dV_j, warpgroup 1 drains dK_j
      const uint32_t src = tmem_base + (which == 0 ? AB_TM_DV : AB_TM_DK) + lane_off;
      const uint32_t dst_row = sDQ_row + which * AB_TILE;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        if (n_it > 0) {
          tmem_ld_32x32(src + c * 32, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;             // no query tile attends to these keys (causal, Sk > Sq)
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const float* f = reinterpret_cast<const float*>(v) + ch * 8;
          st_shared_v4(dst_row + (((c * 4 + ch) ^ sw) << 4), pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                       pack_bf16x2(f[6], f[7]));
        }
      }
    }
    fence_proxy_async_smem();
    named_bar_sync(3, 256);
    if (threadIdx.x == 0) {
      tma_store_3d_b(&tmap_dv, sDQ, h * AB_HD, k0, b);
      tma_store_3d_b(&tmap_dk, sDQ + AB_TILE, h * AB_HD, k0, b);
      tma_store_commit();
      tma_store_wait_all<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// backward preprocess (also used by the mma.sync path): delta[b,h,i] = sum_d dO[i,d] * O[i,d]
__global__ void attn_bwd_delta_tc_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout, float* __restrict__ delta, int64_t ldo,
                                         int64_t lddo, int B, int H, int Sq) {
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp_global >= B * H * Sq) return;
  const int i = warp_global % Sq, bh = warp_global / Sq, h = bh % H, b = bh / H;
  const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(o + ((int64_t)b * Sq + i) * ldo + h * AB_HD + lane * 2));
  const float2 g = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dout + ((int64_t)b * Sq + i) * lddo + h * AB_HD + lane * 2));
  const float s = warp_sum(a.x * g.x + a.y * g.y);
  if (lane == 0) delta[warp_global] = s;
}

// [B, S, cols] view of a [B*S, ld] matrix; box = box_cols x 128 rows x 1 batch, 128B swizzle
static int make_tmap_bsc2(CUtensorMap* out, const void* base, int elem_bytes, int64_t ld, int B, int S, int cols, int box_cols) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) { dwb_set_error("cuTensorMapEncodeTiled entry point unavailable"); return DWB_ERR_CUDA; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((ld * elem_bytes) & 15) != 0) {
    dwb_set_error("attention operand needs a 16 B aligned base and row pitch (base=%p pitch=%lld B)", base, (long long)ld * elem_bytes);
    return DWB_ERR_INVALID;
  }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)S, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * elem_bytes, (cuuint64_t)S * (cuuint64_t)ld * elem_bytes};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base),
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { dwb_set_error("cuTensorMapEncodeTiled(3d) failed with %d", (int)r); return DWB_ERR_CUDA; }
  return DWB_OK;
}

}  // namespace dwb

using namespace dwb;

extern "C" int dwb_attention_bwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* o,
                                    int64_t ldo, const void* dout, int64_t lddo, const float* lse, float* delta_ws, float* dq_acc,
                                    void* dk, int64_t lddk, void* dv, int64_t lddv, int B, int H, int Sq, int Sk, int head_dim, int causal,
                                    float scale, void* stream) {
  DWB_CHECK_ARG(head_dim == AB_HD, "dwb_attention_bwd_tc: head_dim %d unsupported", head_dim);
  DWB_CHECK_ARG(q && k && v && o && dout && lse && delta_ws && dq_acc && dk && dv, "dwb_attention_bwd_tc: null operand");
  DWB_CHECK_ARG(B > 0 && H > 0 && Sq > 0 && Sk > 0, "dwb_attention_bwd_tc: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const int total = B * H * Sq;
  attn_bwd_delta_tc_kernel<<<ceil_div(total, 8), 256, 0, st>>>((const bf16*)o, (const bf16*)dout, delta_ws, ldo, lddo, B, H, Sq);
  DWB_LAUNCH_OK();
  DWB_CUDA_OK(cudaMemsetAsync(dq_acc, 0, (size_t)B * Sq * H * AB_HD * sizeof(float), st));
  CUtensorMap tq, tk, tv, tdo, tdq, tdk, tdv;
  int rc;
  const int cols = H * AB_HD;
  if ((rc = make_tmap_bsc2(&tq, q, 2, ldq, B, Sq, cols, 64))) return rc;
  if ((rc = make_tmap_bsc2(&tk, k, 2, ldk, B, Sk, cols, 64))) return rc;
  if ((rc = make_tmap_bsc2(&tv, v, 2, ldv, B, Sk, cols, 64))) return rc;
  if ((rc = make_tmap_bsc2(&tdo, dout, 2, lddo, B, Sq, cols, 64))) return rc;
  if ((rc = make_tmap_bsc2(&tdq, dq_acc, 4, cols, B, Sq, cols, 32))) return rc;
  if ((rc = make_tmap_bsc2(&tdk, dk, 2, lddk, B, Sk, cols, 64))) return rc;
  if ((rc = make_tmap_bsc2(&tdv, dv, 2, lddv, B, Sk, cols, 64))) return rc;
  static bool attr = false;
  if (!attr) {
    DWB_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM));
    attr = true;
  }
  TcAttnBwdParams p;
  p.H = H; p.Sq = Sq; p.Sk = Sk; p.causal = causal;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse = lse;
  p.delta = delta_ws;
  dim3 grid(ceil_div(Sk, AB_T), B * H);
  attn_bwd_tc_kernel<<<grid, AB_THREADS, AB_SMEM, st>>>(tq, tk, tv, tdo, tdq, tdk, tdv, p);
  DWB_LAUNCH_OK();
  return DWB_OK;
}
