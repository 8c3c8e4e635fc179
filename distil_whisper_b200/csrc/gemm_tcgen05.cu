// bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma (fp32 accumulators in TMEM, double
// buffered) -> epilogue warps (tcgen05.ld, bias / GELU, bf16 or fp32) -> swizzled smem panel -> TMA store /
// TMA reduce-add.  Persistent: one CTA per SM walks a static tile list.
//
// Replaces every nn.Linear / Conv1d-as-GEMM the reference reaches through
// HF:models/whisper/modeling_whisper.py:310-355 (q/k/v/out_proj), :404-407 (fc1/gelu/fc2), :619-620 (conv1/conv2),
// :1081 (proj_out) and their autograd backward (dgrad / wgrad) under ref:training/run_distillation.py:1609.
//
//   C[M,N] = act(alpha * A[M,K] . B[N,K]^T + bias[N])
// Operand storage is described per operand ("major"):
//   K-major  (0): A is [M,K] row-major / B is [N,K] row-major      (reduction dim contiguous)  -- forward, y = x W^T
//   MN-major (1): A is [K,M] row-major / B is [K,N] row-major      (reduction dim strided)     -- dgrad / wgrad
// Both are fed to the tensor core directly (UMMA a_major/b_major bits); nothing is transposed in HBM.
#include "common.cuh"

namespace dwb {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kGemmThreads = 192;            // warps 0-3 epilogue, warp 4 TMA producer, warp 5 MMA issuer
constexpr int kEpiThreads = 128;
constexpr int kPanelBytes = 128 * 128;        // 128 rows x 128 B of staging per buffer (all four epilogue warps)
constexpr int kWarpPanelBytes = 32 * 128;     // one warp's 32 rows x 128 B

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (192 * 1024) / kStageBytes;
  static constexpr int kTmemCols = BN <= 128 ? 256 : 512;   // double-buffered accumulator (allocation must be a power of two)
  static constexpr int kUsedBytes = kStages * kStageBytes + 2 * kPanelBytes + 256 + 2 * BN * 4;
  static constexpr int kSmemBytes = 512 + kUsedBytes;    // 512 B of slack for rounding the base up to 1024 B (checked)
};

struct GemmParams {
  int M, N, K;
  int split_k;          // >1: each work item reduces a K slice and reduce-adds fp32 into C
  int c_f32;            // output dtype: 0 bf16, 1 fp32
  int reduce_add;       // TMA reduce-add instead of store (fp32 only)
  int act;              // 0 none, 1 gelu(erf)
  float alpha;
  const float* bias;    // [N] or null
  int reverse;          // walk the tiles from the last row block to the first (see dwb_set_row_walk)
};

template <int BN, int A_MN, int B_MN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_c, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (threadIdx.x == 0 && (int)(smem - smem_raw) + Cfg::kUsedBytes > Cfg::kSmemBytes) {
    printf("dwb: gemm smem window misaligned by %d B\n", (int)(smem - smem_raw));
    __trap();
  }
  uint8_t* panels = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(panels + 2 * kPanelBytes);
  uint64_t* full_bar = bars;                      // [kStages]
  uint64_t* empty_bar = bars + Cfg::kStages;      // [kStages]
  uint64_t* tmem_full = bars + 2 * Cfg::kStages;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;           // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_bias = reinterpret_cast<float*>(bars) + 64;   // [2][BN] fp32, 256 B past the barriers

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = ceil_div(p.M, BM);
  const int n_tiles = ceil_div(p.N, BN);
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb_total = ceil_div(p.K, BK);
  const int kb_per_split = ceil_div(num_kb_total, p.split_k);
  const int num_items = num_tiles * p.split_k;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_c);
  }
  if (warp == 5 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kEpiThreads);
    }
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc(tmem_ptr, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 4) {
    // ===================================== TMA producer ======================================
    // The whole warp walks the loop (uniform control flow); elect.sync picks the issuing lane, which lets ptxas emit the
    // UTMALDG / UTCHMMA instructions directly instead of wrapping each one in a per-thread election loop.
    {
      const bool leader = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int tile = p.reverse ? num_tiles - 1 - item % num_tiles : item % num_tiles;
        const int split = item / num_tiles;
        const int m0 = (tile / n_tiles) * BM;
        const int n0 = (tile % n_tiles) * BN;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, num_kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) {
            uint8_t* sa = smem + stage * Cfg::kStageBytes;
            uint8_t* sb = sa + Cfg::kABytes;
            mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            if (A_MN) {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma_load_2d(&tmap_a, &full_bar[stage], sa + j * (BK * 128), m0 + 64 * j, kb * BK);
            } else {
              tma_load_2d(&tmap_a, &full_bar[stage], sa, kb * BK, m0);
            }
            if (B_MN) {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j) tma_load_2d(&tmap_b, &full_bar[stage], sb + j * (BK * 128), n0 + 64 * j, kb * BK);
            } else {
              tma_load_2d(&tmap_b, &full_bar[stage], sb, kb * BK, n0);
            }
          }
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ===================================== MMA issuer ========================================
    // Highest warp id in the CTA: the SMSP arbiter favours it over the epilogue warp sharing its sub-partition, so
    // epilogue math never delays tensor-core issue.
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
      // descriptor start-address advance (16 B units) per UMMA_K = 16 elements of K
      constexpr uint32_t a_kstep = A_MN ? (2 * 1024) >> 4 : 32 >> 4;
      constexpr uint32_t b_kstep = B_MN ? (2 * 1024) >> 4 : 32 >> 4;
      constexpr uint32_t a_lbo = A_MN ? BK * 128 : 16;
      constexpr uint32_t b_lbo = B_MN ? BK * 128 : 16;
      const uint64_t da0 = umma_desc_sw128(smem_u32(smem), a_lbo, 1024);
      const uint64_t db0 = umma_desc_sw128(smem_u32(smem) + Cfg::kABytes, b_lbo, 1024);
      int stage = 0;
      uint32_t phase = 0;
      int local_it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++local_it) {
        const int split = item / num_tiles;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, num_kb_total);
        const int acc = local_it & 1;
        const uint32_t acc_phase = (local_it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (leader) {
            const uint64_t da = da0 + (uint64_t)(stage * (Cfg::kStageBytes >> 4));
            const uint64_t db = db0 + (uint64_t)(stage * (Cfg::kStageBytes >> 4));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              tc_mma_ss(tmem_d, da + (uint64_t)(k * a_kstep), db + (uint64_t)(k * b_kstep), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            tc_commit(&empty_bar[stage]);             // frees the smem stage once these MMAs retire
            if (kb == kb1 - 1) tc_commit(&tmem_full[acc]);
          }
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        if (kb1 <= kb0 && leader) tc_commit(&tmem_full[acc]);   // empty K slice (never with sane split_k)
      }
    }
    __syncwarp();
  } else {
    // ===================================== epilogue (warps 0-3) =========================================
    // Each warp drains its own 32 TMEM lanes (rows) through private smem panels and issues its own TMA stores, so the
    // four warps never wait for each other inside a tile; the only CTA-level sync is one named barrier per tile that
    // publishes the staged bias slice.
    const int q = warp;                           // TMEM lane quadrant this warp may read (warp id % 4)
    const int epi_tid = threadIdx.x;
    const int panel_cols = p.c_f32 ? 32 : 64;
    const int panels_per_tile = BN / panel_cols;
    uint8_t* my_panels = panels + q * (2 * kWarpPanelBytes);            // 2 x (32 rows x 128 B)
    const uint32_t row_saddr = smem_u32(my_panels) + lane * 128;
    const int sw = lane & 7;
    int local_it = 0;
    int panel_it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++local_it) {
      const int tile = p.reverse ? num_tiles - 1 - item % num_tiles : item % num_tiles;
      const int m0 = (tile / n_tiles) * BM + q * 32;
      const int n0 = (tile % n_tiles) * BN;
      const int acc = local_it & 1;
      const uint32_t acc_phase = (local_it >> 1) & 1;
      float* sb = s_bias + (local_it & 1) * BN;   // double buffered: tile t+2's writer is behind tile t+1's barrier
      for (int c = epi_tid; c < BN; c += kEpiThreads) sb[c] = (p.bias != nullptr && n0 + c < p.N) ? __ldg(p.bias + n0 + c) : 0.f;
      named_bar_sync(1, kEpiThreads);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      for (int pi = 0; pi < panels_per_tile; ++pi, ++panel_it) {
        const int c0 = pi * panel_cols;
        if (n0 + c0 >= p.N) break;                // fully out-of-range panel (N tail)
        uint32_t v[64];
        tmem_ld_32x32(t_row + c0, v);
        if (!p.c_f32) tmem_ld_32x32(t_row + c0 + 32, v + 32);
        tmem_ld_wait();
        const bool last_panel = (pi == panels_per_tile - 1) || (n0 + c0 + panel_cols >= p.N);
        if (last_panel) {                         // accumulator fully read: hand TMEM back to the MMA warp
          tc_fence_before();
          mbar_arrive(&tmem_empty[acc]);
        }
        const uint32_t buf_off = (panel_it & 1) * kWarpPanelBytes;
        if (lane == 0) tma_store_wait_read<1>();  // this warp's store from two panels ago has left the buffer
        __syncwarp();
        const float4* sb4 = reinterpret_cast<const float4*>(sb + c0);
        const float2 alpha2 = make_float2(p.alpha, p.alpha);   // packed fp32x2 epilogue math (FFMA2): half the FMA instructions
        if (p.c_f32) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {           // 8 x 16 B chunks (4 fp32) per 128 B row
            const float4 bb = sb4[j];
            float2 g0 = __ffma2_rn(make_float2(__uint_as_float(v[j * 4 + 0]), __uint_as_float(v[j * 4 + 1])), alpha2, make_float2(bb.x, bb.y));
            float2 g1 = __ffma2_rn(make_float2(__uint_as_float(v[j * 4 + 2]), __uint_as_float(v[j * 4 + 3])), alpha2, make_float2(bb.z, bb.w));
            if (p.act == 1) { g0 = gelu_erf_fast2(g0); g1 = gelu_erf_fast2(g1); }
            st_shared_v4(row_saddr + buf_off + ((j ^ sw) << 4), __float_as_uint(g0.x), __float_as_uint(g0.y), __float_as_uint(g1.x),
                         __float_as_uint(g1.y));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {           // 8 x 16 B chunks (8 bf16) per 128 B row
            const float4 b0 = sb4[2 * j], b1 = sb4[2 * j + 1];
            float2 g[4];
            g[0] = __ffma2_rn(make_float2(__uint_as_float(v[j * 8 + 0]), __uint_as_float(v[j * 8 + 1])), alpha2, make_float2(b0.x, b0.y));
            g[1] = __ffma2_rn(make_float2(__uint_as_float(v[j * 8 + 2]), __uint_as_float(v[j * 8 + 3])), alpha2, make_float2(b0.z, b0.w));
            g[2] = __ffma2_rn(make_float2(__uint_as_float(v[j * 8 + 4]), __uint_as_float(v[j * 8 + 5])), alpha2, make_float2(b1.x, b1.y));
            g[3] = __ffma2_rn(make_float2(__uint_as_float(v[j * 8 + 6]), __uint_as_float(v[j * 8 + 7])), alpha2, make_float2(b1.z, b1.w));
            if (p.act == 1) {
#pragma unroll
              for (int e = 0; e < 4; ++e) g[e] = gelu_erf_fast2(g[e]);
            }
            st_shared_v4(row_saddr + buf_off + ((j ^ sw) << 4), pack_bf16x2(g[0].x, g[0].y), pack_bf16x2(g[1].x, g[1].y),
                         pack_bf16x2(g[2].x, g[2].y), pack_bf16x2(g[3].x, g[3].y));
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (p.reduce_add) tma_reduce_add_2d(&tmap_c, my_panels + buf_off, n0 + c0, m0);
          else tma_store_2d(&tmap_c, my_panels + buf_off, n0 + c0, m0);
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a cluster of two CTAs (the two SMs of a TPC) computes a 256 x 256 tile with one
// 256-row tcgen05.mma per K step.  Each CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256
// columns); the tensor cores of both SMs read the two B halves from both shared memories.  Per SM and K block that is
// 32 KB of TMA writes and operand reads instead of 48 KB, and the ring holds 6 stages instead of 4.
//   * both producers' TMA bytes are accounted on the LEADER's (cluster rank 0) full barrier; only the leader issues MMAs;
//   * tcgen05.commit multicasts "stage free" / "accumulator ready" to the barriers at the same offset in both CTAs;
//   * the peer's epilogue warps release the accumulator with a remote mbarrier arrive on the leader.
struct Gemm2Cfg {
  static constexpr int BN = 256, BNH = 128;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BNH * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (192 * 1024) / kStageBytes;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kUsedBytes = kStages * kStageBytes + 2 * kPanelBytes + 256 + 2 * BN * 4;
  static constexpr int kSmemBytes = 512 + kUsedBytes;
};
static_assert(2 * Gemm2Cfg::kStages + 4 <= 31, "barriers + TMEM pointer must fit in the 256 B in front of the bias slices");

template <int A_MN, int B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_bf16_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                      const __grid_constant__ CUtensorMap tmap_c, const GemmParams p) {
  using Cfg = Gemm2Cfg;
  constexpr int BN = Cfg::BN, BNH = Cfg::BNH;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (threadIdx.x == 0 && (int)(smem - smem_raw) + Cfg::kUsedBytes > Cfg::kSmemBytes) {
    printf("dwb: gemm smem window misaligned by %d B\n", (int)(smem - smem_raw));
    __trap();
  }
  uint8_t* panels = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(panels + 2 * kPanelBytes);
  uint64_t* full_bar = bars;                      // [kStages]   (only the leader's are waited on)
  uint64_t* empty_bar = bars + Cfg::kStages;      // [kStages]
  uint64_t* tmem_full = bars + 2 * Cfg::kStages;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;           // [2]        (only the leader's are waited on)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_bias = reinterpret_cast<float*>(bars) + 64;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  const int m_tiles = ceil_div(p.M, 2 * BM);
  const int n_tiles = ceil_div(p.N, BN);
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb_total = ceil_div(p.K, BK);
  const int kb_per_split = ceil_div(num_kb_total, p.split_k);
  const int num_items = num_tiles * p.split_k;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_c);
  }
  if (warp == 5 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);               // one arrive per epilogue warp of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc_2cta(tmem_ptr, Cfg::kTmemCols);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();                             // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 4) {
    // ===================================== TMA producer (both CTAs) ==========================
    const bool leader = elect_one();
    const uint32_t full0 = mapa_u32(smem_u32(&full_bar[0]), 0);   // the pair leader's full barriers
    int stage = 0;
    uint32_t phase = 0;
    for (int item = cluster_id; item < num_items; item += num_clusters) {
      const int tile = p.reverse ? num_tiles - 1 - item % num_tiles : item % num_tiles;
      const int split = item / num_tiles;
      const int m0 = (tile / n_tiles) * (2 * BM) + (int)rank * BM;
      const int n0 = (tile % n_tiles) * BN + (int)rank * BNH;     // this CTA's half of the B tile
      const int kb0 = split * kb_per_split;
      const int kb1 = min(kb0 + kb_per_split, num_kb_total);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) {
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          const uint32_t fb = full0 + stage * 8;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);   // both CTAs' bytes land on this barrier
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d_2sm(&tmap_a, fb, sa + j * (BK * 128), m0 + 64 * j, kb * BK);
          } else {
            tma_load_2d_2sm(&tmap_a, fb, sa, kb * BK, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BNH / 64; ++j) tma_load_2d_2sm(&tmap_b, fb, sb + j * (BK * 128), n0 + 64 * j, kb * BK);
          } else {
            tma_load_2d_2sm(&tmap_b, fb, sb, kb * BK, n0);
          }
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ===================================== MMA issuer (leader CTA only) ======================
    if (rank == 0) {
      const bool leader = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN, A_MN, B_MN);
      constexpr uint32_t a_kstep = A_MN ? (2 * 1024) >> 4 : 32 >> 4;
      constexpr uint32_t b_kstep = B_MN ? (2 * 1024) >> 4 : 32 >> 4;
      constexpr uint32_t a_lbo = A_MN ? BK * 128 : 16;
      constexpr uint32_t b_lbo = B_MN ? BK * 128 : 16;
      const uint64_t da0 = umma_desc_sw128(smem_u32(smem), a_lbo, 1024);
      const uint64_t db0 = umma_desc_sw128(smem_u32(smem) + Cfg::kABytes, b_lbo, 1024);
      int stage = 0;
      uint32_t phase = 0;
      int local_it = 0;
      for (int item = cluster_id; item < num_items; item += num_clusters, ++local_it) {
        const int split = item / num_tiles;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, num_kb_total);
        const int acc = local_it & 1;
        const uint32_t acc_phase = (local_it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (leader) {
            const uint64_t da = da0 + (uint64_t)(stage * (Cfg::kStageBytes >> 4));
            const uint64_t db = db0 + (uint64_t)(stage * (Cfg::kStageBytes >> 4));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              tc_mma_ss_2cta(tmem_d, da + (uint64_t)(k * a_kstep), db + (uint64_t)(k * b_kstep), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            tc_commit_2cta(&empty_bar[stage], 3);             // frees the stage in both CTAs
            if (kb == kb1 - 1) tc_commit_2cta(&tmem_full[acc], 3);
          }
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        if (kb1 <= kb0 && leader) tc_commit_2cta(&tmem_full[acc], 3);
      }
    }
    __syncwarp();
  } else {
    // ===================================== epilogue (warps 0-3 of both CTAs) ==================
    const int q = warp;
    const int epi_tid = threadIdx.x;
    const int panel_cols = p.c_f32 ? 32 : 64;
    const int panels_per_tile = BN / panel_cols;
    uint8_t* my_panels = panels + q * (2 * kWarpPanelBytes);
    const uint32_t row_saddr = smem_u32(my_panels) + lane * 128;
    const int sw = lane & 7;
    const uint32_t tmem_empty0 = mapa_u32(smem_u32(&tmem_empty[0]), 0);   // the leader's accumulator-free barriers
    int local_it = 0;
    int panel_it = 0;
    for (int item = cluster_id; item < num_items; item += num_clusters, ++local_it) {
      const int tile = p.reverse ? num_tiles - 1 - item % num_tiles : item % num_tiles;
      const int m0 = (tile / n_tiles) * (2 * BM) + (int)rank * BM + q * 32;
      const int n0 = (tile % n_tiles) * BN;
      const int acc = local_it & 1;
      const uint32_t acc_phase = (local_it >> 1) & 1;
      float* sb = s_bias + (local_it & 1) * BN;
      for (int c = epi_tid; c < BN; c += kEpiThreads) sb[c] = (p.bias != nullptr && n0 + c < p.N) ? __ldg(p.bias + n0 + c) : 0.f;
      named_bar_sync(1, kEpiThreads);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      bool released = false;
      for (int pi = 0; pi < panels_per_tile; ++pi, ++panel_it) {
        const int c0 = pi * panel_cols;
        if (n0 + c0 >= p.N) break;
        uint32_t v[64];
        tmem_ld_32x32(t_row + c0, v);
        if (!p.c_f32) tmem_ld_32x32(t_row + c0 + 32, v + 32);
        tmem_ld_wait();
        const bool last_panel = (pi == panels_per_tile - 1) || (n0 + c0 + panel_cols >= p.N);
        if (last_panel) {                         // accumulator fully read by this warp: tell the leader's MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tmem_empty0 + acc * 8);
          released = true;
        }
        const uint32_t buf_off = (panel_it & 1) * kWarpPanelBytes;
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
        const float4* sb4 = reinterpret_cast<const float4*>(sb + c0);
        const float2 alpha2 = make_float2(p.alpha, p.alpha);   // packed fp32x2 epilogue math (FFMA2): half the FMA instructions
        if (p.c_f32) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = sb4[j];
            float2 g0 = __ffma2_rn(make_float2(__uint_as_float(v[j * 4 + 0]), __uint_as_float(v[j * 4 + 1])), alpha2, make_float2(bb.x, bb.y));
            float2 g1 = __ffma2_rn(make_float2(__uint_as_float(v[j * 4 + 2]), __uint_as_float(v[j * 4 + 3])), alpha2, make_float2(bb.z, bb.w));
            if (p.act == 1) { g0 = gelu_erf_fast2(g0); g1 = gelu_erf_fast2(g1); }
            st_shared_v4(row_saddr + buf_off + ((j ^ sw) << 4), __float_as_uint(g0.x), __float_as_uint(g0.y), __float_as_uint(g1.x),
                         __float_as_uint(g1.y));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b0 = sb4[2 * j], b1 = sb4[2 * j + 1];
            float2 g[4];
            g[0] = __ffma2_rn(make_float2(__uint_as_float(v[j * 8 + 0]), __uint_as_float(v[j * 8 + 1])), alpha2, make_float2(b0.x, b0.y));
            g[1] = __ffma2_rn(make_float2(__uint_as_float(v[j * 8 + 2]), __uint_as_float(v[j * 8 + 3])), alpha2, make_float2(b0.z, b0.w));
            g[2] = __ffma2_rn(make_float2(__uint_as_float(v[j * 8 + 4]), __uint_as_float(v[j * 8 + 5])), alpha2, make_float2(b1.x, b1.y));
            g[3] = __ffma2_rn(make_float2(__uint_as_float(v[j * 8 + 6]), __uint_as_float(v[j * 8 + 7])), alpha2, make_float2(b1.z, b1.w));
            if (p.act == 1) {
#pragma unroll
              for (int e = 0; e < 4; ++e) g[e] = gelu_erf_fast2(g[e]);
            }
            st_shared_v4(row_saddr + buf_off + ((j ^ sw) << 4), pack_bf16x2(g[0].x, g[0].y), pack_bf16x2(g[1].x, g[1].y),
                         pack_bf16x2(g[2].x, g[2].y), pack_bf16x2(g[3].x, g[3].y));
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (p.reduce_add) tma_reduce_add_2d(&tmap_c, my_panels + buf_off, n0 + c0, m0);
          else tma_store_2d(&tmap_c, my_panels + buf_off, n0 + c0, m0);
          tma_store_commit();
        }
      }
      if (!released) {                            // (cannot happen: n0 < N for every scheduled tile) keep the protocol whole
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tmem_empty0 + acc * 8);
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  cluster_sync_all();                             // both CTAs are done with both TMEMs / shared memories
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// plain SIMT GEMM with the same contract: debugging aid and on-device cross-check for the tests.
__global__ void gemm_bf16_simt_kernel(const bf16* __restrict__ A, int64_t lda, int a_mn, const bf16* __restrict__ B,
                                      int64_t ldb, int b_mn, void* __restrict__ C, int64_t ldc, int c_f32, int M, int N,
                                      int K, const float* __restrict__ bias, int act, float alpha, int accumulate) {
  __shared__ float sa[16][17], sb[16][17];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int m = blockIdx.y * 16 + ty, n = blockIdx.x * 16 + tx;
  float acc = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    const int ka = k0 + tx;
    sa[ty][tx] = (m < M && ka < K) ? __bfloat162float(a_mn ? A[(int64_t)ka * lda + m] : A[(int64_t)m * lda + ka]) : 0.f;
    const int nb = blockIdx.x * 16 + ty;
    sb[ty][tx] = (nb < N && ka < K) ? __bfloat162float(b_mn ? B[(int64_t)ka * ldb + nb] : B[(int64_t)nb * ldb + ka]) : 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc += sa[ty][k] * sb[tx][k];
    __syncthreads();
  }
  if (m < M && n < N) {
    float x = acc * alpha;
    if (bias) x += bias[n];
    if (act == 1) x = gelu_erf(x);
    if (c_f32) {
      float* c = reinterpret_cast<float*>(C) + (int64_t)m * ldc + n;
      *c = accumulate ? (*c + x) : x;
    } else {
      reinterpret_cast<bf16*>(C)[(int64_t)m * ldc + n] = __float2bfloat16_rn(x);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
static PFN_encodeTiled g_encode = nullptr;
PFN_encodeTiled get_encode_tiled() {
  if (g_encode) return g_encode;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
  return g_encode;
}

int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                 uint32_t box_rows, uint32_t box_cols) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) { dwb_set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)"); return DWB_ERR_CUDA; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((ld_elems * elem_bytes) & 15) != 0) {
    dwb_set_error("TMA operand needs a 16 B aligned base and row pitch (base=%p, pitch=%llu B)", base,
                  (unsigned long long)(ld_elems * elem_bytes));
    return DWB_ERR_INVALID;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = enc(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    dwb_set_error("cuTensorMapEncodeTiled failed with %d (rows=%llu cols=%llu ld=%llu box=%ux%u)", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, box_rows, box_cols);
    return DWB_ERR_CUDA;
  }
  return DWB_OK;
}

template <int BN, int A_MN, int B_MN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmParams& p, int grid,
                       cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_bf16_tcgen05_kernel<BN, A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    DWB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  kern<<<grid, kGemmThreads, Cfg::kSmemBytes, st>>>(ta, tb, tc, p);
  DWB_LAUNCH_OK();
  return DWB_OK;
}

template <int A_MN, int B_MN>
static int launch_gemm_2cta(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmParams& p, int items,
                            cudaStream_t st) {
  auto kern = gemm_bf16_2cta_kernel<A_MN, B_MN>;
  static int max_clusters = 0;
  if (max_clusters == 0) {
    DWB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Cfg::kSmemBytes));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * kNumSMs);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = Gemm2Cfg::kSmemBytes;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = kNumSMs / 2;
    }
    max_clusters = n;
  }
  const int clusters = items < max_clusters ? items : max_clusters;
  kern<<<2 * clusters, kGemmThreads, Gemm2Cfg::kSmemBytes, st>>>(ta, tb, tc, p);
  DWB_LAUNCH_OK();
  return DWB_OK;
}

static int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || g_num_sms <= 0)
      g_num_sms = kNumSMs;
  }
  return g_num_sms;
}

}  // namespace dwb

using namespace dwb;

extern "C" int dwb_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major,
                             void* C, int64_t ldc, int c_f32, int M, int N, int K, const float* bias, int act,
                             float alpha, int accumulate, int impl, void* stream) {
  DWB_CHECK_ARG(A && B && C, "dwb_gemm_bf16: null operand");
  DWB_CHECK_ARG(M > 0 && N > 0 && K > 0, "dwb_gemm_bf16: bad shape M=%d N=%d K=%d", M, N, K);
  DWB_CHECK_ARG(!(accumulate && !c_f32), "dwb_gemm_bf16: accumulate needs an fp32 C");
  DWB_CHECK_ARG(act == 0 || act == 1, "dwb_gemm_bf16: unknown activation %d", act);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (impl == 1) {   // SIMT cross-check implementation
    dim3 grid(ceil_div(N, 16), ceil_div(M, 16)), block(16, 16);
    gemm_bf16_simt_kernel<<<grid, block, 0, st>>>(reinterpret_cast<const bf16*>(A), lda, a_mn_major,
                                                   reinterpret_cast<const bf16*>(B), ldb, b_mn_major, C, ldc, c_f32, M,
                                                   N, K, bias, act, alpha, accumulate);
    DWB_LAUNCH_OK();
    return DWB_OK;
  }
  DWB_CHECK_ARG(impl == 0 || impl == 2 || impl == 3, "dwb_gemm_bf16: unknown impl %d (0 auto, 1 SIMT, 2 CTA pair, 3 single CTA)", impl);

  // tile shape / split-K heuristic: fill 148 SMs in as few full waves as possible
  const int sms = num_sms();
  const int m_tiles = ceil_div(M, BM);
  // 128x256 tiles feed the tensor pipe at 96 B/clk of smem reads; 128x128 tiles need 128 B/clk (the smem limit) and reach
  // ~70 % of the 256-wide rate, 128x192 sits in between.  Pick the width whose (rounds x cost of one tile) is smallest:
  // e.g. the decoder's 4096x1280 outputs are 160 256-wide tiles = two rounds on 148 SMs, but 224 192-wide tiles = two
  // rounds of 3/4 the work.
  int bn = 256;
  double best_single = 1e30;
  {
    const int cand[3] = {256, 192, 128};
    const double cost[3] = {1.0, 0.78, 0.72};
    double best = 1e30;
    for (int i = 0; i < 3; ++i) {
      if (cand[i] != 128 && N <= 128) continue;
      const double w = (double)ceil_div(m_tiles * ceil_div(N, cand[i]), sms) * cost[i];
      if (w < best - 1e-9) { best = w; bn = cand[i]; }
    }
    best_single = best;
  }
  // CTA pairs (256 x 256 tiles, 74 clusters): worth it when the problem fills the pairs for several rounds
  static const int pair_env = [] { const char* e = getenv("DWB_GEMM_2CTA"); return e ? atoi(e) : 1; }();
  bool pair = false;
  if (impl == 2) pair = true;
  else if (impl == 0 && pair_env && N >= 256) {
    const int t2 = ceil_div(M, 2 * BM) * ceil_div(N, 256), clusters = sms / 2;
    // measured (profiles/r01_kernel_microbench_v5/v6.json): a pair finishes its 256x256 tile 5-12 % sooner than two single
    // CTAs finish their 128x256 tiles once it has >= 8 tiles to walk (M = 48000 shapes, LM head); with fewer tiles the
    // 128x192 single-CTA tiling quantises better (decoder shapes), except for very long reductions (LM-head dgrad)
    const double pair_cost = (double)ceil_div(t2, clusters) * 0.92;
    pair = (t2 >= 8 * clusters && pair_cost <= best_single + 1e-9) || (t2 >= clusters && K >= 16384);
  }
  if (pair) bn = 256;
  const int tiles = pair ? ceil_div(M, 2 * BM) * ceil_div(N, 256) : m_tiles * ceil_div(N, bn);
  const int num_kb = ceil_div(K, BK);
  int split_k = 1;
  if (c_f32 && bias == nullptr && act == 0 && tiles * 2 <= sms && num_kb >= 16) {
    split_k = sms / tiles;
    if (split_k > num_kb / 4) split_k = num_kb / 4;
    if (split_k < 1) split_k = 1;
    // every split must own at least one k block
    while (split_k > 1 && ceil_div(num_kb, split_k) * (split_k - 1) >= num_kb) --split_k;
  }
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.split_k = split_k;
  p.c_f32 = c_f32;
  p.reduce_add = (accumulate || split_k > 1) ? 1 : 0;
  p.act = act;
  p.alpha = alpha;
  p.bias = bias;
  p.reverse = dwb_row_walk_reverse();
  if (split_k > 1 && !accumulate) {
    DWB_CUDA_OK(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, st));
  }

  CUtensorMap ta, tb, tc;
  int rc;
  if (a_mn_major) rc = make_tmap_2d(&ta, A, 2, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, 64);
  else rc = make_tmap_2d(&ta, A, 2, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, BK);
  if (rc) return rc;
  if (b_mn_major) rc = make_tmap_2d(&tb, B, 2, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, 64);
  else rc = make_tmap_2d(&tb, B, 2, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, (uint32_t)(pair ? 128 : bn), BK);
  if (rc) return rc;
  rc = make_tmap_2d(&tc, C, c_f32 ? 4 : 2, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, 32, c_f32 ? 32 : 64);   // one epilogue warp's slab
  if (rc) return rc;

  const int items = tiles * split_k;
  if (pair) {
    switch ((a_mn_major ? 2 : 0) | (b_mn_major ? 1 : 0)) {
      case 0: return launch_gemm_2cta<0, 0>(ta, tb, tc, p, items, st);
      case 1: return launch_gemm_2cta<0, 1>(ta, tb, tc, p, items, st);
      case 2: return launch_gemm_2cta<1, 0>(ta, tb, tc, p, items, st);
      default: return launch_gemm_2cta<1, 1>(ta, tb, tc, p, items, st);
    }
  }
  const int grid = items < sms ? items : sms;
  const int key = (bn == 256 ? 4 : bn == 192 ? 8 : 0) | (a_mn_major ? 2 : 0) | (b_mn_major ? 1 : 0);
  switch (key) {
    case 8: return launch_gemm<192, 0, 0>(ta, tb, tc, p, grid, st);
    case 9: return launch_gemm<192, 0, 1>(ta, tb, tc, p, grid, st);
    case 10: return launch_gemm<192, 1, 0>(ta, tb, tc, p, grid, st);
    case 11: return launch_gemm<192, 1, 1>(ta, tb, tc, p, grid, st);
    case 0: return launch_gemm<128, 0, 0>(ta, tb, tc, p, grid, st);
    case 1: return launch_gemm<128, 0, 1>(ta, tb, tc, p, grid, st);
    case 2: return launch_gemm<128, 1, 0>(ta, tb, tc, p, grid, st);
    case 3: return launch_gemm<128, 1, 1>(ta, tb, tc, p, grid, st);
    case 4: return launch_gemm<256, 0, 0>(ta, tb, tc, p, grid, st);
    case 5: return launch_gemm<256, 0, 1>(ta, tb, tc, p, grid, st);
    case 6: return launch_gemm<256, 1, 0>(ta, tb, tc, p, grid, st);
    default: return launch_gemm<256, 1, 1>(ta, tb, tc, p, grid, st);
  }
}
