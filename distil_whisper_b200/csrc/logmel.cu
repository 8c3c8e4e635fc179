// Whisper log-mel feature extractor on sm_100a.
//   wav [B, 480000] fp32  ->  out [B, n_mels, 3000] fp32
// Restates HF:models/whisper/feature_extraction_whisper.py:135-164 (_torch_extract_fbank_features): hann(400)
// periodic window, STFT n_fft 400 / hop 160 / center + reflect padding, drop the last frame, |.|^2, slaney mel bank,
// log10(max(., 1e-10)), clamp to (per-utterance max - 8), (x + 4) / 4.
//
// Design (HBM-bound target: 1.92 MB read + 0.96 MB write per utterance, nothing else):
//   * one 8-CTA thread-block cluster per utterance; CTA r owns frames [376 r, 376 r + 376)
//   * samples are staged once in shared memory (coalesced loads, reflect padding resolved at load time); the 2.5x
//     frame overlap is served from smem, not HBM
//   * 400-point real FFT = 200-point complex FFT (8 x 5 x 5 mixed radix, all stages in shared memory) + split
//   * the CTA's whole log-mel slab stays in shared memory until the per-utterance max is known: CTA maxima are
//     exchanged through distributed shared memory (cluster.map_shared_rank), so the output is written exactly once
//     and never re-read -- no second pass over HBM for the `max - 8` floor.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace dwb {

constexpr int LM_NFFT = 400;
constexpr int LM_HOP = 160;
constexpr int LM_NFREQ = 201;
constexpr int LM_CLUSTER = 8;
constexpr int LM_THREADS = 512;
constexpr int LM_POW_PITCH = 203;

struct LogmelPlan {
  int n_mels;
  int max_w;                 // widest filter (bins)
  float* d_window;           // [400]
  float2* d_tw200;           // [200] exp(-2 pi i j / 200)
  float2* d_tw25;            // [25]
  float2* d_tw400;           // [201] exp(-2 pi i k / 400)
  int* d_mel_start;          // [n_mels]
  int* d_mel_count;          // [n_mels]
  float* d_mel_w;            // [n_mels, max_w]
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

// forward 4-point DFT in place
__device__ __forceinline__ void dft4(float2& c0, float2& c1, float2& c2, float2& c3) {
  const float2 t0 = cadd(c0, c2), t1 = csub(c0, c2), t2 = cadd(c1, c3), t3 = mul_mi(csub(c1, c3));
  c0 = cadd(t0, t2); c2 = csub(t0, t2); c1 = cadd(t1, t3); c3 = csub(t1, t3);
}
// forward 5-point DFT in place
__device__ __forceinline__ void dft5(float2& x0, float2& x1, float2& x2, float2& x3, float2& x4) {
  const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f, s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
  const float2 t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
  const float2 m1 = make_float2(x0.x + c1 * t1.x + c2 * t2.x, x0.y + c1 * t1.y + c2 * t2.y);
  const float2 m2 = make_float2(x0.x + c2 * t1.x + c1 * t2.x, x0.y + c2 * t1.y + c1 * t2.y);
  const float2 n1 = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
  const float2 n2 = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
  x0 = make_float2(x0.x + t1.x + t2.x, x0.y + t1.y + t2.y);
  // X1 = m1 - i n1, X4 = m1 + i n1, X2 = m2 - i n2, X3 = m2 + i n2
  x1 = make_float2(m1.x + n1.y, m1.y - n1.x);
  x4 = make_float2(m1.x - n1.y, m1.y + n1.x);
  x2 = make_float2(m2.x + n2.y, m2.y - n2.x);
  x3 = make_float2(m2.x - n2.y, m2.y + n2.x);
}

__global__ void __launch_bounds__(LM_THREADS, 1)
logmel_kernel(const LogmelPlan plan, const float* __restrict__ wav, float* __restrict__ out, int n_samples, int n_frames,
              int frames_per_cta, int F) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int n_mels = plan.n_mels;
  // carve
  float* out_s = reinterpret_cast<float*>(smem_raw);                               // [n_mels][frames_per_cta]
  float* s_window = out_s + (size_t)n_mels * frames_per_cta;                       // [400]
  float2* s_tw200 = reinterpret_cast<float2*>(s_window + LM_NFFT);                 // [200]
  float2* s_tw25 = s_tw200 + 200;                                                  // [25] (+1 pad)
  float2* s_tw400 = s_tw25 + 26;                                                   // [201] (+1 pad)
  float* s_melw = reinterpret_cast<float*>(s_tw400 + 202);                         // [n_mels][wp]  (wp odd: conflict-free)
  const int wp = plan.max_w | 1;
  float* s_samples = s_melw + (((size_t)n_mels * wp + 3) & ~(size_t)3);            // [(F-1)*160 + 400]
  const int span_max = (F - 1) * LM_HOP + LM_NFFT;
  float2* bufA = reinterpret_cast<float2*>(s_samples + span_max);                  // [F][200]
  float2* bufB = bufA + (size_t)F * 200;                                           // [F][200]  (aliased by pow [F][203])
  __shared__ float s_red[LM_THREADS / 32];
  __shared__ float s_cta_max;

  const int tid = threadIdx.x;
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.y;
  const float* w = wav + (size_t)b * n_samples;
  const int f_begin = rank * frames_per_cta;
  const int f_end = min(n_frames, f_begin + frames_per_cta);

  for (int i = tid; i < LM_NFFT; i += LM_THREADS) s_window[i] = plan.d_window[i];
  for (int i = tid; i < 200; i += LM_THREADS) s_tw200[i] = plan.d_tw200[i];
  for (int i = tid; i < 25; i += LM_THREADS) s_tw25[i] = plan.d_tw25[i];
  for (int i = tid; i < LM_NFREQ; i += LM_THREADS) s_tw400[i] = plan.d_tw400[i];
  for (int i = tid; i < n_mels * plan.max_w; i += LM_THREADS) s_melw[(i / plan.max_w) * wp + (i % plan.max_w)] = plan.d_mel_w[i];
  // this lane's mel filters (m = lane + 32 i): first bin and width, kept in registers for the whole kernel
  int mel_st[4], mel_cnt[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = (tid & 31) + 32 * i;
    mel_st[i] = m < n_mels ? __ldg(plan.d_mel_start + m) : 0;
    mel_cnt[i] = m < n_mels ? __ldg(plan.d_mel_count + m) : 0;
  }

  float lmax = -INFINITY;
  for (int f0 = f_begin; f0 < f_end; f0 += F) {
    const int nf = min(F, f_end - f0);
    __syncthreads();   // previous chunk fully consumed (and tables visible on the first trip)
    // ---- 1. stage the samples (reflect padding of n_fft/2 on both ends, torch.stft center=True) ----
    const int span = (nf - 1) * LM_HOP + LM_NFFT;
    const int p0 = f0 * LM_HOP - LM_NFFT / 2;
    for (int i = tid; i < span; i += LM_THREADS) {
      int o = p0 + i;
      if (o < 0) o = -o;
      if (o >= n_samples) o = 2 * (n_samples - 1) - o;
      s_samples[i] = __ldg(w + o);
    }
    __syncthreads();
    // ---- per-frame pipeline: each warp owns whole frames (f = warp, warp+16, ...), stages separated by __syncwarp only ----
    const int warp = tid >> 5, lane = tid & 31;
    for (int f = warp; f < nf; f += LM_THREADS / 32) {
      float2* zA = bufA + (size_t)f * 200;
      float2* zB = bufB + (size_t)f * 200;
      float* pw = reinterpret_cast<float*>(zB);           // power spectrum of this frame (after stage 4 has consumed zB)
      // -- 2. radix-8 over n1 (z[25 n1 + n2]), twiddle W200^(n2 k1): lanes 0..24
      if (lane < 25) {
        const int n2 = lane;
        const float* xs = s_samples + f * LM_HOP + 2 * n2;
        const float* ws = s_window + 2 * n2;
        float2 x[8];
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
          const float2 sv = *reinterpret_cast<const float2*>(xs + 50 * n1);
          const float2 wv = *reinterpret_cast<const float2*>(ws + 50 * n1);
          x[n1] = make_float2(sv.x * wv.x, sv.y * wv.y);
        }
        float2 a0 = cadd(x[0], x[4]), a1 = cadd(x[1], x[5]), a2 = cadd(x[2], x[6]), a3 = cadd(x[3], x[7]);
        float2 b0 = csub(x[0], x[4]), b1 = csub(x[1], x[5]), b2 = csub(x[2], x[6]), b3 = csub(x[3], x[7]);
        const float r = 0.70710678118654752f;
        b1 = make_float2((b1.x + b1.y) * r, (b1.y - b1.x) * r);      // * W8^1 = (1 - i)/sqrt2
        b2 = mul_mi(b2);                                              // * W8^2 = -i
        b3 = make_float2((b3.y - b3.x) * r, -(b3.x + b3.y) * r);     // * W8^3 = (-1 - i)/sqrt2
        dft4(a0, a1, a2, a3);   // X[0], X[2], X[4], X[6]
        dft4(b0, b1, b2, b3);   // X[1], X[3], X[5], X[7]
        float2* dst = zA + n2;
        dst[0] = a0;
        dst[25] = cmul(b0, s_tw200[n2]);
        dst[50] = cmul(a1, s_tw200[2 * n2]);
        dst[75] = cmul(b1, s_tw200[3 * n2]);
        dst[100] = cmul(a2, s_tw200[4 * n2]);
        dst[125] = cmul(b2, s_tw200[5 * n2]);
        dst[150] = cmul(a3, s_tw200[6 * n2]);
        dst[175] = cmul(b3, s_tw200[7 * n2]);
      }
      __syncwarp();
      // -- 3. DFT-25 = 5 x 5: radix-5 over a (n2 = 5a + b), twiddle W25^(b c); 40 items = lanes 0..31 then 0..7
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const int r = pass * 32 + lane;
        if (r < 40) {
          const int k1 = r / 5, bb = r - k1 * 5;
          const float2* src = zA + k1 * 25 + bb;
          float2 y0 = src[0], y1 = src[5], y2 = src[10], y3 = src[15], y4 = src[20];
          dft5(y0, y1, y2, y3, y4);
          float2* dst = zB + k1 * 25 + bb * 5;
          dst[0] = y0;
          dst[1] = cmul(y1, s_tw25[bb]);
          dst[2] = cmul(y2, s_tw25[2 * bb]);
          dst[3] = cmul(y3, s_tw25[3 * bb]);
          dst[4] = cmul(y4, s_tw25[4 * bb]);
        }
      }
      __syncwarp();
      // -- 4. second radix-5 over b -> Z[k1 + 8 (c + 5 e)]
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const int r = pass * 32 + lane;
        if (r < 40) {
          const int k1 = r / 5, c = r - k1 * 5;
          const float2* src = zB + k1 * 25 + c;
          float2 u0 = src[0], u1 = src[5], u2 = src[10], u3 = src[15], u4 = src[20];
          dft5(u0, u1, u2, u3, u4);
          float2* dst = zA + k1 + 8 * c;
          dst[0] = u0; dst[40] = u1; dst[80] = u2; dst[120] = u3; dst[160] = u4;
        }
      }
      __syncwarp();
      // -- 5. real-FFT split + power spectrum, bins k and 200-k together (X[k] = E + W^k O, conj X[200-k] = E - W^k O)
#pragma unroll
      for (int pass = 0; pass < 4; ++pass) {
        const int k = pass * 32 + lane;                    // 0..100
        if (k <= 100) {
          const float2 zk = zA[k];
          float2 zr = zA[k == 0 ? 0 : 200 - k];
          zr.y = -zr.y;
          const float2 e = make_float2(0.5f * (zk.x + zr.x), 0.5f * (zk.y + zr.y));
          const float2 d = make_float2(0.5f * (zk.x - zr.x), 0.5f * (zk.y - zr.y));
          const float2 t = cmul(mul_mi(d), s_tw400[k]);
          const float re0 = e.x + t.x, im0 = e.y + t.y, re1 = e.x - t.x, im1 = e.y - t.y;
          pw[k] = re0 * re0 + im0 * im0;
          pw[200 - k] = re1 * re1 + im1 * im1;             // k = 100 writes the same bin twice with the same value
        }
      }
      __syncwarp();
      // -- 6. mel projection + log10: lanes over mel bins
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = lane + 32 * i;
        if (m < n_mels) {
          const float* wrow = s_melw + m * wp;
          const float* prow = pw + mel_st[i];
          float acc0 = 0.f, acc1 = 0.f;
          int j = 0;
          for (; j + 1 < mel_cnt[i]; j += 2) {
            acc0 = fmaf(wrow[j], prow[j], acc0);
            acc1 = fmaf(wrow[j + 1], prow[j + 1], acc1);
          }
          if (j < mel_cnt[i]) acc0 = fmaf(wrow[j], prow[j], acc0);
          const float lv = log10f(fmaxf(acc0 + acc1, 1e-10f));
          out_s[(size_t)m * frames_per_cta + (f0 - f_begin) + f] = lv;
          lmax = fmaxf(lmax, lv);
        }
      }
      __syncwarp();
    }
  }
  // ---- per-utterance max through distributed shared memory ----
  lmax = warp_max(lmax);
  if ((tid & 31) == 0) s_red[tid >> 5] = lmax;
  __syncthreads();
  if (tid < 32) {
    float v = tid < LM_THREADS / 32 ? s_red[tid] : -INFINITY;
    v = warp_max(v);
    if (tid == 0) s_cta_max = v;
  }
  cluster.sync();
  float umax = -INFINITY;
#pragma unroll
  for (int r = 0; r < LM_CLUSTER; ++r) umax = fmaxf(umax, *cluster.map_shared_rank(&s_cta_max, r));
  cluster.sync();     // nobody leaves (and frees its smem) before all peers have read it
  const float floor_v = umax - 8.0f;
  const int nfl = f_end - f_begin;
  float* ob = out + (size_t)b * n_mels * n_frames + f_begin;
  for (int it = tid; it < n_mels * nfl; it += LM_THREADS) {
    const int m = it / nfl, f = it - m * nfl;
    const float v = fmaxf(out_s[(size_t)m * frames_per_cta + f], floor_v);
    ob[(size_t)m * n_frames + f] = (v + 4.0f) * 0.25f;
  }
}

}  // namespace dwb

using namespace dwb;

extern "C" int dwb_logmel_plan_destroy(void* plan_v) {
  if (!plan_v) return DWB_OK;
  LogmelPlan* p = reinterpret_cast<LogmelPlan*>(plan_v);
  cudaFree(p->d_window); cudaFree(p->d_tw200); cudaFree(p->d_tw25); cudaFree(p->d_tw400);
  cudaFree(p->d_mel_start); cudaFree(p->d_mel_count); cudaFree(p->d_mel_w);
  free(p);
  return DWB_OK;
}

extern "C" int dwb_logmel_plan_create(const float* mel_filters_host, int n_freq, int n_mels, void** plan_out) {
  DWB_CHECK_ARG(mel_filters_host && plan_out, "dwb_logmel_plan_create: null argument");
  DWB_CHECK_ARG(n_freq == LM_NFREQ, "dwb_logmel_plan_create: expected %d frequency bins (n_fft 400), got %d", LM_NFREQ, n_freq);
  DWB_CHECK_ARG(n_mels > 0 && n_mels <= 128, "dwb_logmel_plan_create: n_mels=%d unsupported (1..128)", n_mels);
  const double PI = 3.14159265358979323846;
  float window[LM_NFFT];
  float2 tw200[200], tw25[25], tw400[LM_NFREQ];
  for (int i = 0; i < LM_NFFT; ++i) window[i] = (float)(0.5 - 0.5 * cos(2.0 * PI * i / LM_NFFT));   // periodic hann
  for (int i = 0; i < 200; ++i) tw200[i] = make_float2((float)cos(2.0 * PI * i / 200.0), (float)-sin(2.0 * PI * i / 200.0));
  for (int i = 0; i < 25; ++i) tw25[i] = make_float2((float)cos(2.0 * PI * i / 25.0), (float)-sin(2.0 * PI * i / 25.0));
  for (int i = 0; i < LM_NFREQ; ++i) tw400[i] = make_float2((float)cos(2.0 * PI * i / 400.0), (float)-sin(2.0 * PI * i / 400.0));
  // sparse view of the [n_freq, n_mels] bank: each triangular filter is a contiguous run of bins
  int* start = (int*)malloc(sizeof(int) * n_mels);
  int* count = (int*)malloc(sizeof(int) * n_mels);
  int max_w = 1;
  for (int m = 0; m < n_mels; ++m) {
    int lo = -1, hi = -1;
    for (int k = 0; k < n_freq; ++k)
      if (mel_filters_host[(size_t)k * n_mels + m] != 0.f) { if (lo < 0) lo = k; hi = k; }
    start[m] = lo < 0 ? 0 : lo;
    count[m] = lo < 0 ? 0 : hi - lo + 1;
    if (count[m] > max_w) max_w = count[m];
  }
  float* wts = (float*)calloc((size_t)n_mels * max_w, sizeof(float));
  for (int m = 0; m < n_mels; ++m)
    for (int j = 0; j < count[m]; ++j) wts[(size_t)m * max_w + j] = mel_filters_host[(size_t)(start[m] + j) * n_mels + m];

  LogmelPlan* p = (LogmelPlan*)calloc(1, sizeof(LogmelPlan));
  p->n_mels = n_mels;
  p->max_w = max_w;
  cudaError_t e = cudaSuccess;
#define LM_UP(dst, src, bytes)                                                     \
  if (e == cudaSuccess) e = cudaMalloc((void**)&(dst), (bytes));                   \
  if (e == cudaSuccess) e = cudaMemcpy((dst), (src), (bytes), cudaMemcpyHostToDevice);
  LM_UP(p->d_window, window, sizeof(window));
  LM_UP(p->d_tw200, tw200, sizeof(tw200));
  LM_UP(p->d_tw25, tw25, sizeof(tw25));
  LM_UP(p->d_tw400, tw400, sizeof(tw400));
  LM_UP(p->d_mel_start, start, sizeof(int) * n_mels);
  LM_UP(p->d_mel_count, count, sizeof(int) * n_mels);
  LM_UP(p->d_mel_w, wts, sizeof(float) * (size_t)n_mels * max_w);
#undef LM_UP
  free(start); free(count); free(wts);
  if (e != cudaSuccess) {
    dwb_set_error("dwb_logmel_plan_create: %s", cudaGetErrorString(e));
    dwb_logmel_plan_destroy(p);
    return DWB_ERR_CUDA;
  }
  *plan_out = p;
  return DWB_OK;
}

extern "C" int dwb_logmel(void* plan_v, const float* wav, int B, int n_samples, float* out, void* stream) {
  DWB_CHECK_ARG(plan_v && wav && out, "dwb_logmel: null argument");
  DWB_CHECK_ARG(B > 0 && n_samples >= LM_NFFT && (n_samples % LM_HOP) == 0, "dwb_logmel: n_samples=%d must be a positive multiple of %d",
                n_samples, LM_HOP);
  const LogmelPlan* p = reinterpret_cast<const LogmelPlan*>(plan_v);
  const int n_frames = n_samples / LM_HOP;
  int fpc = ceil_div(n_frames, LM_CLUSTER);
  fpc = (fpc + 3) & ~3;
  const size_t fixed = (size_t)p->n_mels * fpc * 4 + LM_NFFT * 4 + (200 + 26 + 202) * 8 + (LM_NFFT - LM_HOP) * 4 + 256 +
                       ((size_t)p->n_mels * (p->max_w | 1) + 4) * 4;
  const size_t per_frame = LM_HOP * 4 + 2 * 200 * 8;   // samples + bufA + bufB (pow aliases bufB: 203*4 <= 1600)
  const size_t budget = 227 * 1024;
  DWB_CHECK_ARG(fixed + per_frame <= budget, "dwb_logmel: slab for n_mels=%d x %d frames does not fit in shared memory", p->n_mels, fpc);
  int F = (int)((budget - fixed) / per_frame);
  if (F > 32) F = 32;
  const size_t smem = fixed + per_frame * F;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    DWB_CUDA_OK(cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(LM_CLUSTER, B, 1);
  cfg.blockDim = dim3(LM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = LM_CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DWB_CUDA_OK(cudaLaunchKernelEx(&cfg, logmel_kernel, *p, wav, out, n_samples, n_frames, fpc, F));
  return DWB_OK;
}
