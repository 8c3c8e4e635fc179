// Whisper log-mel on the tensor cores (sm_100a): the windowed 400-point DFT of every frame as a tcgen05 GEMM.
//   wav [B, 480000] fp32  ->  out [B, n_mels, 3000] fp32          (same contract and arithmetic as logmel.cu / dwb_logmel;
//   HF:models/whisper/feature_extraction_whisper.py:135-164)
//
// Why a GEMM: the shared-memory FFT of logmel.cu is instruction-issue bound (5.5 % of the HBM roofline, profiles/r01_ncu_logmel.md);
// frames x [hann * cos | -hann * sin] is 2 x 400 x 400 MACs per frame, 0.98 TFLOP per 1024 clips -- noise for tcgen05 -- provided
// fp32 accuracy survives.  It does with a two-term fp16 split of both operands and three MMAs per k-step:
//     x = x_hi + x_lo,  W = W_hi + W_lo  (each term 11 significant bits)      X = x_hi W_hi + x_hi W_lo + x_lo W_hi  (+ O(2^-22))
// accumulated in fp32 in TMEM.  The waveform is pre-scaled by 64 so that x_lo stays above fp16's subnormal floor for every
// sample that matters (|x| >= 2^-9; below that the absolute error is 5e-10), and 64^-2 is folded into the mel weights.
//
// Pipeline per chunk of utterances (sized so that tiles = a multiple of 148 and the fp16 scratch stays L2-resident):
//   1. logmel_split_kernel : wav -> x_hi, x_lo fp16 with the reflect padding of torch.stft(center=True) materialised
//   2. logmel_dft_kernel   : persistent, one CTA per SM, 128-frame tiles.  The A operand (frames x 400 samples) is NOT built:
//        a 3-D tensor map with a frame stride of 160 samples (320 B) over the padded fp16 signal lets TMA deliver the overlapping
//        windows straight into the 128B-swizzled K-major operand layout.  B = [hann cos | -hann sin] tables (fp16 hi / lo) stream
//        from L2.  Two passes (real, imaginary) x 7 k-blocks of 64 through a 2-stage smem ring; accumulators Re, Im = 2 x 208 TMEM
//        columns.  Epilogue: thread = frame (TMEM lane), power = re^2 + im^2, the sparse slaney bank (every bin feeds at most two
//        adjacent triangles, bins in increasing order -> two running registers, each finished mel bin is written once, coalesced
//        over frames), log10 via lg2, per-utterance maximum by an ordered-int atomicMax.
//   3. logmel_finish_kernel: max(x, umax - 8), (x + 4) / 4 in place (L2-resident chunk).
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace dwb {

constexpr int LT_NFFT = 400, LT_HOP = 160, LT_NFREQ = 201;
constexpr int LT_KPAD = 448;               // 7 k-blocks of 64 (columns >= 400 of the tables are zero)
constexpr int LT_NB = 208;                 // accumulator width (201 bins padded to a multiple of 16)
constexpr int LT_BM = 128, LT_BK = 64;
constexpr int LT_KB = LT_KPAD / LT_BK;     // 7
// The 208 bins are computed in two halves (112 + 96 bins) with their own Re / Im accumulators (2 x 112 + 2 x 96 = 416 TMEM columns):
// while the epilogue warps turn one half into mel energies, the MMA warp is already accumulating the other half (or the next tile).
constexpr int LT_N0 = 112, LT_N1 = 96;
// One smem stage = one (half, k-block): A hi / lo (the 128 frames x 64 samples window tile, loaded once) + the real and the
// imaginary table tiles hi / lo of that half -- six MMAs per k-step share the A tile, which keeps the L2 -> SM operand traffic
// (the resource this kernel is bound by) at 1.19 MB per 128-frame tile.
constexpr int LT_STAGES = 2;
constexpr int LT_A_BYTES = LT_BM * LT_BK * 2;          // 16384
constexpr int LT_B_BYTES = LT_N0 * LT_BK * 2;          // 14336 (room for the wider half)
constexpr int LT_STAGE_BYTES = 2 * LT_A_BYTES + 4 * LT_B_BYTES;   // 90112
constexpr int LT_THREADS = 192;            // warps 0-3 epilogue, 4 TMA, 5 MMA
constexpr int LT_TMEM_COLS = 512;
constexpr int LT_SMEM = LT_STAGES * LT_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 3 * LT_NB * 4 /*bank tables*/;
constexpr float LT_PRESCALE = 64.f;

struct LogmelTcPlan {
  int n_mels;
  __half* d_w_hi;            // [416, LT_KPAD] K-major; rows: bins 0..111 real (cos) | bins 0..111 imaginary (-sin) | bins 112..207 real | imaginary
  __half* d_w_lo;
  int* d_bin_m0;             // [LT_NB] number of mel filters completed before bin k is accumulated ("rotations"), >= 0
  float* d_bin_w0;           // [LT_NB] weight into the current filter (x 64^-2; 0 for bins that feed nothing)
  float* d_bin_w1;           // [LT_NB] weight into the next filter
  // chunk pipeline (owned by the plan: one dwb_logmel_tc call at a time per plan): the split / finish kernels of neighbouring chunks
  // run on `mem` while the DFT kernel of the current chunk runs on `mma`; both are forked from and joined back into the caller's stream
  cudaStream_t mem, mma;
  cudaEvent_t ev_fork, ev_split[2], ev_dft[2], ev_join_mem, ev_join_mma;
};

// padded signal length per utterance: n_samples + 400 (reflect padding), rounded up to a multiple of the hop so that the utterance
// stride of the window view is an integer multiple of its frame stride (a tensor-map requirement)
__host__ __device__ inline int64_t lt_padded_len(int n_samples) { return ((int64_t)n_samples + LT_NFFT + LT_HOP - 1) / LT_HOP * LT_HOP; }

// ---- 1. split ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) logmel_split_kernel(const float* __restrict__ wav, __half* __restrict__ hi, __half* __restrict__ lo,
                                                           int n_samples, int64_t lp) {
  const int u = blockIdx.y;
  const float* w = wav + (size_t)u * n_samples;
  __half* h = hi + (size_t)u * lp;
  __half* l = lo + (size_t)u * lp;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < lp; i += (int64_t)gridDim.x * blockDim.x * 2) {
    float x[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      int64_t o = i + e - LT_NFFT / 2;                     // reflect padding (no edge repeat), torch.stft center=True
      if (o < 0) o = -o;
      if (o >= n_samples) o = 2 * ((int64_t)n_samples - 1) - o;
      x[e] = __ldg(w + o) * LT_PRESCALE;
    }
    const __half2 hh = __floats2half2_rn(x[0], x[1]);
    const float2 hf = __half22float2(hh);
    *reinterpret_cast<__half2*>(h + i) = hh;
    *reinterpret_cast<__half2*>(l + i) = __floats2half2_rn(x[0] - hf.x, x[1] - hf.y);
  }
}

// ---- 2. DFT GEMM + mel + log ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ordered_int(float f) {
  const int b = __float_as_int(f);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float from_ordered_int(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

constexpr uint32_t lt_idesc_f16(int M, int N) {       // kind::f16, A = B = fp16 (format 0), D = fp32, both K-major
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(LT_THREADS, 1)
logmel_dft_kernel(const __grid_constant__ CUtensorMap tmap_a_hi, const __grid_constant__ CUtensorMap tmap_a_lo,
                  const __grid_constant__ CUtensorMap tmap_w_hi0, const __grid_constant__ CUtensorMap tmap_w_lo0,
                  const __grid_constant__ CUtensorMap tmap_w_hi1, const __grid_constant__ CUtensorMap tmap_w_lo1, const LogmelTcPlan plan,
                  float* __restrict__ out, int* __restrict__ umax_ord, int n_utt, int n_frames, int tiles_per_utt, int dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LT_STAGES * LT_STAGE_BYTES);
  uint64_t* full_bar = bars;                 // [LT_STAGES <= 3]
  uint64_t* empty_bar = bars + 3;            // [LT_STAGES <= 3]
  uint64_t* acc_full = bars + 6;             // [2] one per half
  uint64_t* acc_empty = bars + 8;            // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 10);
  int* s_m0 = reinterpret_cast<int*>(bars + 32);          // 256 B past the barrier block
  float* s_w0 = reinterpret_cast<float*>(s_m0 + LT_NB);
  float* s_w1 = s_w0 + LT_NB;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = n_utt * tiles_per_utt;

  for (int i = threadIdx.x; i < LT_NB; i += LT_THREADS) {
    s_m0[i] = plan.d_bin_m0[i];
    s_w0[i] = plan.d_bin_w0[i];
    s_w1[i] = plan.d_bin_w1[i];
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmap_a_hi); tma_prefetch_desc(&tmap_a_lo); tma_prefetch_desc(&tmap_w_hi0); tma_prefetch_desc(&tmap_w_lo0);
    tma_prefetch_desc(&tmap_w_hi1); tma_prefetch_desc(&tmap_w_lo1);
  }
  if (warp == 5 && lane == 0) {
    for (int i = 0; i < LT_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc(tmem_ptr, LT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 4) {
    // ===================================== TMA producer ======================================
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int u = tile / tiles_per_utt, f0 = (tile % tiles_per_utt) * LT_BM;
      for (int it = 0; it < 2 * LT_KB; ++it) {
        const int half = it / LT_KB, kb = it % LT_KB;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader && (dbg & 4)) {
          mbar_arrive(&full_bar[stage]);                   // timing experiment: no operand traffic
        } else if (leader) {
          uint8_t* s = smem + stage * LT_STAGE_BYTES;
          const int nb = half ? LT_N1 : LT_N0;
          mbar_expect_tx(&full_bar[stage], 2 * LT_A_BYTES + 4 * nb * LT_BK * 2);
          // frames f0.. x samples [64 kb, 64 kb + 64): rows of the overlapping-window view (frame stride 160 samples)
          tma_load_3d(&tmap_a_hi, &full_bar[stage], s, kb * LT_BK, f0, u);
          tma_load_3d(&tmap_a_lo, &full_bar[stage], s + LT_A_BYTES, kb * LT_BK, f0, u);
#pragma unroll
          for (int pass = 0; pass < 2; ++pass) {
            const int row = half ? 2 * LT_N0 + pass * LT_N1 : pass * LT_N0;
            uint8_t* sb = s + 2 * LT_A_BYTES + pass * 2 * LT_B_BYTES;
            tma_load_2d(half ? &tmap_w_hi1 : &tmap_w_hi0, &full_bar[stage], sb, kb * LT_BK, row);
            tma_load_2d(half ? &tmap_w_lo1 : &tmap_w_lo0, &full_bar[stage], sb + LT_B_BYTES, kb * LT_BK, row);
          }
        }
        __syncwarp();
        if (++stage == LT_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 5) {
    // ===================================== MMA issuer ========================================
    const bool leader = elect_one();
    constexpr uint32_t idesc0 = lt_idesc_f16(LT_BM, LT_N0), idesc1 = lt_idesc_f16(LT_BM, LT_N1);
    int stage = 0;
    uint32_t phase = 0;
    int local_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local_it) {
      for (int it = 0; it < 2 * LT_KB; ++it) {
        const int half = it / LT_KB, kb = it % LT_KB;
        if (kb == 0) {
          mbar_wait(&acc_empty[half], (local_it & 1) ^ 1);   // the epilogue has drained this half's Re / Im of the previous tile
          tc_fence_after();
        }
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (leader) {
          const uint32_t sa = smem_u32(smem + stage * LT_STAGE_BYTES);
          const uint64_t a_hi = umma_desc_sw128(sa, 16, 1024), a_lo = umma_desc_sw128(sa + LT_A_BYTES, 16, 1024);
          const uint32_t idesc = half ? idesc1 : idesc0;
          const int nks = (dbg & 2) ? 0 : kb == LT_KB - 1 ? (LT_NFFT - (LT_KB - 1) * LT_BK) / 16 : LT_BK / 16;      // last k-block: samples 384..399 only
#pragma unroll
          for (int pass = 0; pass < 2; ++pass) {
            const uint32_t sb = sa + 2 * LT_A_BYTES + pass * 2 * LT_B_BYTES;
            const uint64_t b_hi = umma_desc_sw128(sb, 16, 1024), b_lo = umma_desc_sw128(sb + LT_B_BYTES, 16, 1024);
            const uint32_t d = tmem_base + (half ? 2 * LT_N0 + pass * LT_N1 : pass * LT_N0);
#pragma unroll 1
            for (int k = 0; k < nks; ++k) {
              const uint64_t o = (uint64_t)(2 * k);
              tc_mma_ss(d, a_lo + o, b_hi + o, idesc, (kb > 0 || k > 0) ? 1u : 0u);      // small terms first
              tc_mma_ss(d, a_hi + o, b_lo + o, idesc, 1u);
              tc_mma_ss(d, a_hi + o, b_hi + o, idesc, 1u);
            }
          }
          tc_commit(&empty_bar[stage]);
          if (kb == LT_KB - 1) tc_commit(&acc_full[half]);
        }
        __syncwarp();
        if (++stage == LT_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================================== epilogue: thread = frame ==========================
    const int n_mels = plan.n_mels;
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    int local_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local_it) {
      const int u = tile / tiles_per_utt, f = (tile % tiles_per_utt) * LT_BM + warp * 32 + lane;
      const bool live = f < n_frames;
      float* orow = out + (size_t)u * n_mels * n_frames + f;
      if (dbg & 1) {                                  // timing experiment: no epilogue arithmetic
        for (int half = 0; half < 2; ++half) {
          mbar_wait(&acc_full[half], local_it & 1);
          tc_fence_after();
          tc_fence_before();
          mbar_arrive(&acc_empty[half]);
        }
        continue;
      }
      // Two running sums: `a` = the filter being completed, `b` = the next one (every bin feeds at most these two).  The plan
      // stores, per bin, how many filters are complete before it ("rotations") and its two weights; the per-chunk metadata is
      // fetched with a few warp-uniform 16 B shared-memory loads up front, so the unrolled bin loop is FMAs plus a uniform
      // branch, and every completed filter costs one lg2 and one coalesced store.
      float a = 0.f, b = 0.f, vmax = -INFINITY;
      int m_cur = 0;
      float* optr = orow;
      auto rotate = [&]() {
        if (m_cur < n_mels) {
          float l2;                                    // a >= 1e-10: never denormal, the plain approximation is exact enough
          asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(fmaxf(a, 1e-10f)));
          const float lv = l2 * 0.30102999566398120f;
          if (live) *optr = lv;
          vmax = fmaxf(vmax, lv);
        }
        a = b; b = 0.f; ++m_cur; optr += n_frames;
      };
#pragma unroll 1
      for (int c0 = 0; c0 < LT_NB; c0 += 16) {
        const int half = c0 >= LT_N0;
        if (c0 == 0 || c0 == LT_N0) {
          mbar_wait(&acc_full[half], local_it & 1);
          tc_fence_after();
        }
        const uint32_t t_re = t_lane + (half ? 2 * LT_N0 + (c0 - LT_N0) : c0);
        const uint32_t t_im = t_re + (half ? LT_N1 : LT_N0);
        uint32_t re[16], im[16];
        tmem_ld_32x16(t_re, re);
        tmem_ld_32x16(t_im, im);
        float w0[16], w1[16];
        int rot[16];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 x = *reinterpret_cast<const float4*>(s_w0 + c0 + i);
          const float4 y = *reinterpret_cast<const float4*>(s_w1 + c0 + i);
          const int4 z = *reinterpret_cast<const int4*>(s_m0 + c0 + i);
          w0[i] = x.x; w0[i + 1] = x.y; w0[i + 2] = x.z; w0[i + 3] = x.w;
          w1[i] = y.x; w1[i + 1] = y.y; w1[i + 2] = y.z; w1[i + 3] = y.w;
          rot[i] = z.x; rot[i + 1] = z.y; rot[i + 2] = z.z; rot[i + 3] = z.w;
        }
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (rot[i] > 0) {                                 // warp-uniform (the same bin for every frame)
            rotate();
#pragma unroll 1
            for (int r = rot[i]; r > 1; --r) rotate();      // more than one filter ending at once: only where the bank is denser than the bins
          }
          const float rr = __uint_as_float(re[i]), q = __uint_as_float(im[i]);
          const float pw = fmaf(rr, rr, q * q);
          a = fmaf(w0[i], pw, a);
          b = fmaf(w1[i], pw, b);
        }
        if (c0 + 16 == LT_N0 || c0 + 16 == LT_NB) {
          tc_fence_before();
          mbar_arrive(&acc_empty[half]);              // this half's Re / Im are consumed: the MMA warp may overwrite them
        }
      }
#pragma unroll 1
      while (m_cur < n_mels) rotate();
      vmax = live ? vmax : -INFINITY;
      vmax = warp_max(vmax);
      if (lane == 0 && vmax > -INFINITY) atomicMax(umax_ord + u, ordered_int(vmax));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, LT_TMEM_COLS);
  }
}

// ---- 3. floor at (utterance max - 8), (x + 4) / 4 ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) logmel_finish_kernel(float* __restrict__ out, const int* __restrict__ umax_ord, int64_t per_utt) {
  const int u = blockIdx.y;
  const float floor_v = from_ordered_int(umax_ord[u]) - 8.0f;
  float4* o = reinterpret_cast<float4*>(out + (size_t)u * per_utt);
  const int64_t n4 = per_utt >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = o[i];
    v.x = (fmaxf(v.x, floor_v) + 4.0f) * 0.25f; v.y = (fmaxf(v.y, floor_v) + 4.0f) * 0.25f;
    v.z = (fmaxf(v.z, floor_v) + 4.0f) * 0.25f; v.w = (fmaxf(v.w, floor_v) + 4.0f) * 0.25f;
    o[i] = v;
  }
}
__global__ void logmel_init_umax_kernel(int* __restrict__ umax_ord, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) umax_ord[i] = ordered_int(-INFINITY);
}

static int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                         const cuuint32_t* box) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) { dwb_set_error("cuTensorMapEncodeTiled entry point unavailable"); return DWB_ERR_CUDA; }
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { dwb_set_error("cuTensorMapEncodeTiled(fp16, rank %d) failed with %d", rank, (int)r); return DWB_ERR_CUDA; }
  return DWB_OK;
}

constexpr int LT_CHUNK = 37;     // utterances per chunk: 37 x 24 tiles = 6 x 148, and 71 MB of fp16 scratch + 36 MB of output stay in L2

}  // namespace dwb

using namespace dwb;

extern "C" int dwb_logmel_tc_plan_destroy(void* plan_v) {
  if (!plan_v) return DWB_OK;
  LogmelTcPlan* p = reinterpret_cast<LogmelTcPlan*>(plan_v);
  cudaFree(p->d_w_hi); cudaFree(p->d_w_lo); cudaFree(p->d_bin_m0); cudaFree(p->d_bin_w0); cudaFree(p->d_bin_w1);
  if (p->mem) cudaStreamDestroy(p->mem);
  if (p->mma) cudaStreamDestroy(p->mma);
  cudaEvent_t* evs[] = {&p->ev_fork, &p->ev_split[0], &p->ev_split[1], &p->ev_dft[0], &p->ev_dft[1], &p->ev_join_mem, &p->ev_join_mma};
  for (cudaEvent_t* e : evs) if (*e) cudaEventDestroy(*e);
  free(p);
  return DWB_OK;
}

extern "C" int dwb_logmel_tc_plan_create(const float* mel_filters_host, int n_freq, int n_mels, void** plan_out) {
  DWB_CHECK_ARG(mel_filters_host && plan_out, "dwb_logmel_tc_plan_create: null argument");
  DWB_CHECK_ARG(n_freq == LT_NFREQ, "dwb_logmel_tc_plan_create: expected %d frequency bins (n_fft 400), got %d", LT_NFREQ, n_freq);
  DWB_CHECK_ARG(n_mels > 0 && n_mels <= 128, "dwb_logmel_tc_plan_create: n_mels=%d unsupported (1..128)", n_mels);
  const double PI = 3.14159265358979323846;
  const size_t wn = (size_t)2 * LT_NB * LT_KPAD;
  __half* w_hi = (__half*)calloc(wn, sizeof(__half));
  __half* w_lo = (__half*)calloc(wn, sizeof(__half));
  for (int part = 0; part < 2; ++part)
    for (int k = 0; k < LT_NFREQ; ++k)
      for (int n = 0; n < LT_NFFT; ++n) {
        const double win = 0.5 - 0.5 * cos(2.0 * PI * n / LT_NFFT);                       // periodic hann
        const long ph = ((long)k * n) % LT_NFFT;                                         // exact argument reduction
        const double val = win * (part == 0 ? cos(2.0 * PI * ph / LT_NFFT) : -sin(2.0 * PI * ph / LT_NFFT));
        const __half h = __float2half_rn((float)val);
        const size_t row = k < LT_N0 ? (size_t)part * LT_N0 + k : (size_t)2 * LT_N0 + (size_t)part * LT_N1 + (k - LT_N0);
        const size_t idx = row * LT_KPAD + n;
        w_hi[idx] = h;
        w_lo[idx] = __float2half_rn((float)(val - (double)__half2float(h)));
      }
  // sparse slaney bank: every frequency bin feeds at most two adjacent triangular filters
  int m0[LT_NB];
  float w0[LT_NB], w1[LT_NB];
  const float inv = 1.0f / (LT_PRESCALE * LT_PRESCALE);
  int rc = DWB_OK;
  for (int k = 0; k < LT_NB; ++k) {
    m0[k] = -1; w0[k] = 0.f; w1[k] = 0.f;
    if (k >= LT_NFREQ) continue;
    int first = -1, last = -1;
    for (int m = 0; m < n_mels; ++m)
      if (mel_filters_host[(size_t)k * n_mels + m] != 0.f) { if (first < 0) first = m; last = m; }
    if (first < 0) continue;
    if (last - first > 1) { dwb_set_error("dwb_logmel_tc_plan_create: bin %d feeds filters %d..%d (not a two-triangle bank)", k, first, last); rc = DWB_ERR_UNSUPPORTED; break; }
    m0[k] = first;
    w0[k] = mel_filters_host[(size_t)k * n_mels + first] * inv;
    w1[k] = last > first ? mel_filters_host[(size_t)k * n_mels + last] * inv : 0.f;
  }
  // the epilogue walks the bins in increasing order with two running filters: first-filter indices must not decrease; convert
  // them to "filters completed before this bin" counts
  int prev = 0;
  for (int k = 0; k < LT_NB && rc == DWB_OK; ++k) {
    if (m0[k] < 0) { m0[k] = 0; continue; }          // bin feeds nothing: no rotation, zero weights
    if (m0[k] < prev) { dwb_set_error("dwb_logmel_tc_plan_create: mel bank is not ordered by frequency at bin %d", k); rc = DWB_ERR_UNSUPPORTED; break; }
    const int first = m0[k];
    m0[k] = first - prev;
    prev = first;
  }
  if (rc != DWB_OK) { free(w_hi); free(w_lo); return rc; }
  LogmelTcPlan* p = (LogmelTcPlan*)calloc(1, sizeof(LogmelTcPlan));
  p->n_mels = n_mels;
  cudaError_t e = cudaSuccess;
#define LT_UP(dst, src, bytes)                                                     \
  if (e == cudaSuccess) e = cudaMalloc((void**)&(dst), (bytes));                   \
  if (e == cudaSuccess) e = cudaMemcpy((dst), (src), (bytes), cudaMemcpyHostToDevice);
  LT_UP(p->d_w_hi, w_hi, wn * sizeof(__half));
  LT_UP(p->d_w_lo, w_lo, wn * sizeof(__half));
  LT_UP(p->d_bin_m0, m0, sizeof(m0));
  LT_UP(p->d_bin_w0, w0, sizeof(w0));
  LT_UP(p->d_bin_w1, w1, sizeof(w1));
#undef LT_UP
  free(w_hi); free(w_lo);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->mem, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->mma, cudaStreamNonBlocking);
  {
    cudaEvent_t* evs[] = {&p->ev_fork, &p->ev_split[0], &p->ev_split[1], &p->ev_dft[0], &p->ev_dft[1], &p->ev_join_mem, &p->ev_join_mma};
    for (cudaEvent_t* ev : evs) if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    dwb_set_error("dwb_logmel_tc_plan_create: %s", cudaGetErrorString(e));
    dwb_logmel_tc_plan_destroy(p);
    return DWB_ERR_CUDA;
  }
  *plan_out = p;
  return DWB_OK;
}

// workspace: two buffers (double-buffered chunks) of fp16 hi / lo copies of one chunk of padded waveforms + one ordered-int
// maximum per utterance
extern "C" int64_t dwb_logmel_tc_workspace_bytes(int B, int n_samples) {
  const int chunk = B < LT_CHUNK ? B : LT_CHUNK;
  const int64_t lp = lt_padded_len(n_samples);
  const int nbuf = B > chunk ? 2 : 1;
  return nbuf * 2 * (int64_t)chunk * lp * (int64_t)sizeof(__half) + (((int64_t)B * 4 + 255) & ~(int64_t)255) + 256;
}

extern "C" int dwb_logmel_tc(void* plan_v, const float* wav, int B, int n_samples, float* out, void* workspace, void* stream) {
  DWB_CHECK_ARG(plan_v && wav && out && workspace, "dwb_logmel_tc: null argument");
  DWB_CHECK_ARG(B > 0 && n_samples >= LT_NFFT && (n_samples % LT_HOP) == 0 && (n_samples % 8) == 0,
                "dwb_logmel_tc: n_samples=%d must be a positive multiple of %d (and of 8)", n_samples, LT_HOP);
  DWB_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "dwb_logmel_tc: workspace must be 256 B aligned, out 16 B aligned");
  const LogmelTcPlan* p = reinterpret_cast<const LogmelTcPlan*>(plan_v);
  cudaStream_t st = (cudaStream_t)stream;
  const int n_frames = n_samples / LT_HOP;
  DWB_CHECK_ARG(((int64_t)p->n_mels * n_frames) % 4 == 0, "dwb_logmel_tc: n_mels * n_frames must be a multiple of 4");
  const int64_t lp = lt_padded_len(n_samples);
  const int chunk = B < LT_CHUNK ? B : LT_CHUNK;
  int* umax = reinterpret_cast<int*>(workspace);
  __half* scratch = reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(workspace) + ((((int64_t)B * 4 + 255) & ~(int64_t)255)));
  const int64_t buf_elems = 2 * (int64_t)chunk * lp;          // hi + lo of one chunk
  const int tiles_per_utt = ceil_div(n_frames, LT_BM);
  static bool attr = false;
  if (!attr) {
    DWB_CUDA_OK(cudaFuncSetAttribute(logmel_dft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_SMEM));
    attr = true;
  }
  CUtensorMap tw_hi0, tw_lo0, tw_hi1, tw_lo1;
  int rc;
  {
    // tables: [416 rows, 448 cols] K-major; box = 64 cols x 112 rows (first half) or 96 rows (second half)
    const cuuint64_t dims[2] = {LT_KPAD, 2 * LT_NB};
    const cuuint64_t strides[1] = {LT_KPAD * 2};
    const cuuint32_t box0[2] = {LT_BK, LT_N0}, box1[2] = {LT_BK, LT_N1};
    if ((rc = make_tmap_f16(&tw_hi0, p->d_w_hi, 2, dims, strides, box0))) return rc;
    if ((rc = make_tmap_f16(&tw_lo0, p->d_w_lo, 2, dims, strides, box0))) return rc;
    if ((rc = make_tmap_f16(&tw_hi1, p->d_w_hi, 2, dims, strides, box1))) return rc;
    if ((rc = make_tmap_f16(&tw_lo1, p->d_w_lo, 2, dims, strides, box1))) return rc;
  }
  static const int dbg = [] { const char* e = getenv("DWB_LOGMEL_DBG"); return e ? atoi(e) : 0; }();
  const int n_chunks = ceil_div(B, chunk);
  // fork: both internal streams start after everything the caller has queued on `st`
  DWB_CUDA_OK(cudaEventRecord(p->ev_fork, st));
  DWB_CUDA_OK(cudaStreamWaitEvent(p->mem, p->ev_fork, 0));
  DWB_CUDA_OK(cudaStreamWaitEvent(p->mma, p->ev_fork, 0));
  logmel_init_umax_kernel<<<ceil_div(B, 256), 256, 0, p->mem>>>(umax, B);
  DWB_LAUNCH_OK();
  auto issue_split = [&](int c) -> int {
    const int u0 = c * chunk, nu = B - u0 < chunk ? B - u0 : chunk, buf = c & 1;
    __half* hi = scratch + buf * buf_elems;
    if (c >= 2) DWB_CUDA_OK(cudaStreamWaitEvent(p->mem, p->ev_dft[buf], 0));      // the DFT of chunk c-2 has finished reading this buffer
    logmel_split_kernel<<<dim3(64, nu), 256, 0, p->mem>>>(wav + (size_t)u0 * n_samples, hi, hi + (int64_t)chunk * lp, n_samples, lp);
    DWB_LAUNCH_OK();
    DWB_CUDA_OK(cudaEventRecord(p->ev_split[buf], p->mem));
    return DWB_OK;
  };
  if ((rc = issue_split(0))) return rc;
  for (int c = 0; c < n_chunks; ++c) {
    const int u0 = c * chunk, nu = B - u0 < chunk ? B - u0 : chunk, buf = c & 1;
    __half* hi = scratch + buf * buf_elems;
    __half* lo = hi + (int64_t)chunk * lp;
    CUtensorMap ta_hi, ta_lo;
    {
      // overlapping-window view of the padded fp16 signal: [utterance][frame][sample-in-window], frame stride = 160 samples
      const cuuint64_t dims[3] = {LT_KPAD, (cuuint64_t)n_frames, (cuuint64_t)nu};
      const cuuint64_t strides[2] = {LT_HOP * 2, (cuuint64_t)lp * 2};
      const cuuint32_t box[3] = {LT_BK, LT_BM, 1};
      if ((rc = make_tmap_f16(&ta_hi, hi, 3, dims, strides, box))) return rc;
      if ((rc = make_tmap_f16(&ta_lo, lo, 3, dims, strides, box))) return rc;
    }
    DWB_CUDA_OK(cudaStreamWaitEvent(p->mma, p->ev_split[buf], 0));
    const int tiles = nu * tiles_per_utt;
    logmel_dft_kernel<<<tiles < kNumSMs ? tiles : kNumSMs, LT_THREADS, LT_SMEM, p->mma>>>(ta_hi, ta_lo, tw_hi0, tw_lo0, tw_hi1, tw_lo1, *p,
                                                                                          out + (size_t)u0 * p->n_mels * n_frames, umax + u0, nu,
                                                                                          n_frames, tiles_per_utt, dbg);
    DWB_LAUNCH_OK();
    DWB_CUDA_OK(cudaEventRecord(p->ev_dft[buf], p->mma));
    // while that runs: the next chunk's split, then (once the DFT is done) this chunk's finishing pass
    if (c + 1 < n_chunks && (rc = issue_split(c + 1))) return rc;
    DWB_CUDA_OK(cudaStreamWaitEvent(p->mem, p->ev_dft[buf], 0));
    logmel_finish_kernel<<<dim3(64, nu), 256, 0, p->mem>>>(out + (size_t)u0 * p->n_mels * n_frames, umax + u0, (int64_t)p->n_mels * n_frames);
    DWB_LAUNCH_OK();
  }
  // join
  DWB_CUDA_OK(cudaEventRecord(p->ev_join_mem, p->mem));
  DWB_CUDA_OK(cudaEventRecord(p->ev_join_mma, p->mma));
  DWB_CUDA_OK(cudaStreamWaitEvent(st, p->ev_join_mem, 0));
  DWB_CUDA_OK(cudaStreamWaitEvent(st, p->ev_join_mma, 0));
  return DWB_OK;
}
