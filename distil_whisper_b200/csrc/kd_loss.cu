// KD loss head: CE(student, labels) + temperature KL(teacher || student), masked by labels >= 0, and the gradient
// wrt the student logits, in two streaming passes over the [rows, vocab] logits (online softmax in pass 1).
//
// Restates ref:training/run_distillation.py:1453-1462 (kl_divergence), :1484-1493 (CE from the student forward,
// softmax(t/T), log_softmax(s/T), * T^2, loss = 0.8 * ce + kl_weight * kl) and HF:models/whisper/modeling_whisper.py
// :1085-1088 (CrossEntropyLoss, mean over labels != -100) without materialising any [rows, vocab] probability
// tensor: only the bf16 gradient is written.
//
//   ce  = sum_valid(lse(s) - s[label]) / n_valid
//   kl  = T^2 * sum_valid( sum_v p_t (log p_t - log p_s) ) / n_valid,  p_t = softmax(t/T), log p_s = log_softmax(s/T)
//   dloss/ds_v = [ ce_w (softmax(s)_v - 1[v = label]) + kl_w T (softmax(s/T)_v - p_t,v) ] / n_valid   (valid rows)
#include "common.cuh"

namespace dwb {

struct RowStat {
  float ms, z1, zT;   // student: running max, sum exp(s - ms), sum exp((s - ms)/T)
  float mt, zt, acc;  // teacher: running max, sum exp((t - mt)/T), sum exp((t - mt)/T) * (t - s)/T
};

__device__ __forceinline__ void stat_init(RowStat& r) {
  r.ms = -INFINITY; r.z1 = 0.f; r.zT = 0.f; r.mt = -INFINITY; r.zt = 0.f; r.acc = 0.f;
}
__device__ __forceinline__ void stat_push(RowStat& r, float s, float t, float invT, bool has_t) {
  if (s > r.ms) {
    const float d = r.ms - s;                      // <= 0 (or -inf on the first element)
    r.z1 *= __expf(d);
    r.zT *= __expf(d * invT);
    r.ms = s;
  }
  const float es = s - r.ms;
  r.z1 += __expf(es);
  r.zT += __expf(es * invT);
  if (has_t) {
    if (t > r.mt) {
      const float c = __expf((r.mt - t) * invT);
      r.zt *= c;
      r.acc *= c;
      r.mt = t;
    }
    const float e = __expf((t - r.mt) * invT);
    r.zt += e;
    r.acc += e * (t - s) * invT;
  }
}
__device__ __forceinline__ void stat_merge(RowStat& a, const RowStat& b, float invT) {
  const float ms = fmaxf(a.ms, b.ms);
  if (ms != -INFINITY) {
    const float ca = __expf(a.ms - ms), cb = __expf(b.ms - ms);
    const float caT = __expf((a.ms - ms) * invT), cbT = __expf((b.ms - ms) * invT);
    a.z1 = a.z1 * ca + b.z1 * cb;
    a.zT = a.zT * caT + b.zT * cbT;
    a.ms = ms;
  }
  const float mt = fmaxf(a.mt, b.mt);
  if (mt != -INFINITY) {
    const float ca = __expf((a.mt - mt) * invT), cb = __expf((b.mt - mt) * invT);
    a.zt = a.zt * ca + b.zt * cb;
    a.acc = a.acc * ca + b.acc * cb;
    a.mt = mt;
  }
}
__device__ __forceinline__ RowStat stat_shfl_xor(const RowStat& r, int o) {
  RowStat x;
  x.ms = __shfl_xor_sync(0xffffffffu, r.ms, o); x.z1 = __shfl_xor_sync(0xffffffffu, r.z1, o);
  x.zT = __shfl_xor_sync(0xffffffffu, r.zT, o); x.mt = __shfl_xor_sync(0xffffffffu, r.mt, o);
  x.zt = __shfl_xor_sync(0xffffffffu, r.zt, o); x.acc = __shfl_xor_sync(0xffffffffu, r.acc, o);
  return x;
}

__global__ void count_valid_kernel(const int64_t* __restrict__ labels, int rows, int* __restrict__ n_valid) {
  int c = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x) c += labels[i] >= 0 ? 1 : 0;
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(n_valid, c);
}

constexpr int KD_THREADS = 512;

// one CTA per row
__global__ void __launch_bounds__(KD_THREADS) kd_loss_kernel(const float* __restrict__ s_logits, const float* __restrict__ t_logits,
                                                             int64_t ld, const int64_t* __restrict__ labels,
                                                             const int* __restrict__ n_valid_ptr, int V, float temperature,
                                                             float ce_weight, float kl_weight, float* __restrict__ row_ce,
                                                             float* __restrict__ row_kl, bf16* __restrict__ dlogits,
                                                             int64_t ldd) {
  __shared__ RowStat s_red[KD_THREADS / 32];
  __shared__ RowStat s_final;
  const int row = blockIdx.x;
  const int64_t label = labels[row];
  const bool valid = label >= 0;
  const bool has_t = t_logits != nullptr;
  const float invT = 1.f / temperature;
  const int tid = threadIdx.x;
  if (!valid) {
    if (tid == 0) { row_ce[row] = 0.f; row_kl[row] = 0.f; }
    if (dlogits) {
      bf16* d = dlogits + (int64_t)row * ldd;
      for (int c = tid * 8; c < ldd; c += KD_THREADS * 8) *reinterpret_cast<uint4*>(d + c) = make_uint4(0, 0, 0, 0);
    }
    return;
  }
  const float* sr = s_logits + (int64_t)row * ld;
  const float* tr = has_t ? t_logits + (int64_t)row * ld : nullptr;
  RowStat st;
  stat_init(st);
  const int nvec = (V + 3) >> 2;
  for (int i = tid; i < nvec; i += KD_THREADS) {
    const float4 a = *reinterpret_cast<const float4*>(sr + i * 4);
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_t) b = *reinterpret_cast<const float4*>(tr + i * 4);
    const int c = i * 4;
    stat_push(st, a.x, b.x, invT, has_t);
    if (c + 1 < V) stat_push(st, a.y, b.y, invT, has_t);
    if (c + 2 < V) stat_push(st, a.z, b.z, invT, has_t);
    if (c + 3 < V) stat_push(st, a.w, b.w, invT, has_t);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const RowStat other = stat_shfl_xor(st, o);
    stat_merge(st, other, invT);
  }
  if ((tid & 31) == 0) s_red[tid >> 5] = st;
  __syncthreads();
  if (tid < 32) {
    RowStat r;
    stat_init(r);
    if (tid < KD_THREADS / 32) r = s_red[tid];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const RowStat other = stat_shfl_xor(r, o);
      stat_merge(r, other, invT);
    }
    if (tid == 0) s_final = r;
  }
  __syncthreads();
  const RowStat f = s_final;
  const float lse1 = f.ms + __logf(f.z1);                 // log sum exp(s)
  const float lseT = f.ms * invT + __logf(f.zT);          // log sum exp(s/T)
  const float lset = has_t ? f.mt * invT + __logf(f.zt) : 0.f;
  if (tid == 0) {
    // label >= V: CrossEntropyLoss raises (device assert) in the reference; here the row's CE is NaN, so the loss is -- no
    // out-of-bounds read
    row_ce[row] = label < V ? lse1 - sr[label] : __int_as_float(0x7fc00000);
    row_kl[row] = has_t ? (f.acc / f.zt - lset + lseT) : 0.f;
  }
  if (dlogits == nullptr) return;
  const float inv_n = 1.f / (float)max(*n_valid_ptr, 1);
  const float a_ce = ce_weight * inv_n, a_kl = has_t ? kl_weight * temperature * inv_n : 0.f;
  bf16* d = dlogits + (int64_t)row * ldd;
  const int nvec8 = (int)(ldd >> 3);
  for (int i = tid; i < nvec8; i += KD_THREADS) {
    const int c0 = i * 8;
    float g[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (c0 + h * 4 < V) {
        a = *reinterpret_cast<const float4*>(sr + c0 + h * 4);
        if (has_t) b = *reinterpret_cast<const float4*>(tr + c0 + h * 4);
      }
      const float sv[4] = {a.x, a.y, a.z, a.w}, tv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = c0 + h * 4 + e;
        float v = 0.f;
        if (c < V) {
          v = a_ce * (__expf(sv[e] - lse1) - (c == label ? 1.f : 0.f));
          if (has_t) v += a_kl * (__expf(sv[e] * invT - lseT) - __expf(tv[e] * invT - lset));
        }
        g[h * 4 + e] = v;
      }
    }
    *reinterpret_cast<uint4*>(d + c0) =
        make_uint4(pack_bf16x2(g[0], g[1]), pack_bf16x2(g[2], g[3]), pack_bf16x2(g[4], g[5]), pack_bf16x2(g[6], g[7]));
  }
}

// metrics[0..3] = loss, ce, kl, n_valid
__global__ void kd_finalize_kernel(const float* __restrict__ row_ce, const float* __restrict__ row_kl, const int* __restrict__ n_valid,
                                   int rows, float temperature, float ce_weight, float kl_weight, float* __restrict__ metrics) {
  __shared__ double sce[32], skl[32];
  double ce = 0.0, kl = 0.0;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) { ce += row_ce[i]; kl += row_kl[i]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { ce += __shfl_xor_sync(0xffffffffu, ce, o); kl += __shfl_xor_sync(0xffffffffu, kl, o); }
  if ((threadIdx.x & 31) == 0) { sce[threadIdx.x >> 5] = ce; skl[threadIdx.x >> 5] = kl; }
  __syncthreads();
  if (threadIdx.x == 0) {
    ce = 0.0; kl = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { ce += sce[w]; kl += skl[w]; }
    const int n = *n_valid;
    // n == 0 reproduces the reference's 0/0 -> nan (CrossEntropyLoss mean over an empty set)
    const float cef = (float)(ce / (double)n), klf = (float)(kl / (double)n) * temperature * temperature;
    metrics[0] = ce_weight * cef + kl_weight * klf;
    metrics[1] = cef;
    metrics[2] = klf;
    metrics[3] = (float)n;
  }
}

}  // namespace dwb

using namespace dwb;

// workspace: int n_valid (16 B slot) + row_ce[rows] + row_kl[rows]
extern "C" int64_t dwb_kd_loss_workspace_bytes(int rows) { return 16 + (int64_t)rows * 2 * (int64_t)sizeof(float); }

extern "C" int dwb_kd_loss(const float* student_logits, const float* teacher_logits, int64_t ld, const int64_t* labels, int rows,
                           int vocab, float temperature, float ce_weight, float kl_weight, float* metrics4, void* dlogits_bf16,
                           int64_t ldd, void* workspace, void* stream) {
  DWB_CHECK_ARG(student_logits && labels && metrics4 && workspace, "dwb_kd_loss: null operand");
  DWB_CHECK_ARG(rows > 0 && vocab > 0 && (ld % 4) == 0 && ld >= ((vocab + 3) & ~3), "dwb_kd_loss: logits pitch %lld too small / unaligned for vocab %d",
                (long long)ld, vocab);
  DWB_CHECK_ARG(dlogits_bf16 == nullptr || ((ldd % 8) == 0 && ldd >= vocab), "dwb_kd_loss: dlogits pitch");
  DWB_CHECK_ARG(temperature > 0.f, "dwb_kd_loss: temperature must be > 0");
  cudaStream_t st = (cudaStream_t)stream;
  int* n_valid = reinterpret_cast<int*>(workspace);
  float* row_ce = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + 16);
  float* row_kl = row_ce + rows;
  DWB_CUDA_OK(cudaMemsetAsync(n_valid, 0, sizeof(int), st));
  count_valid_kernel<<<ceil_div(rows, 256) < 64 ? ceil_div(rows, 256) : 64, 256, 0, st>>>(labels, rows, n_valid);
  DWB_LAUNCH_OK();
  kd_loss_kernel<<<rows, KD_THREADS, 0, st>>>(student_logits, teacher_logits, ld, labels, n_valid, vocab, temperature, ce_weight,
                                              kl_weight, row_ce, row_kl, (bf16*)dlogits_bf16, ldd);
  DWB_LAUNCH_OK();
  kd_finalize_kernel<<<1, 1024, 0, st>>>(row_ce, row_kl, n_valid, rows, temperature, ce_weight, kl_weight, metrics4);
  DWB_LAUNCH_OK();
  return DWB_OK;
}
