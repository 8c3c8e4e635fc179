// Optimiser tail of the KD step over flat fp32 buffers: gradient sum of squares (for clip_grad_norm_),
// AdamW with the clip coefficient and bias corrections applied on the fly, and the bf16 shadow re-cast that
// the next forward's GEMMs read -- one pass over p/g/m/v.
// Restates ref:training/run_distillation.py:1610-1614 (clip_grad_norm_(max_grad_norm); optimizer.step();
// zero_grad) with torch.optim.AdamW semantics (decoupled weight decay, ref :1402-1407).
#include "common.cuh"

namespace dwb {

// 4 independent 16 B loads per thread and trip: a few dozen CTAs keep enough bytes in flight to stream at HBM speed, which is
// what lets the optimiser tail run with a small grid underneath the next step's encoder forward (dwb_set_tail_grid)
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ out) {
  float acc = 0.f;
  const int64_t nvec = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < nvec; i += 4 * stride) {
    float4 a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = reinterpret_cast<const float4*>(g)[i + u * stride];
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += a[u].x * a[u].x + a[u].y * a[u].y + a[u].z * a[u].z + a[u].w * a[u].w;
  }
  for (; i < nvec; i += stride) {
    const float4 a = reinterpret_cast<const float4*>(g)[i];
    acc += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
  }
  for (int64_t j = (nvec << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) acc += g[j] * g[j];
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = red[threadIdx.x];
    v += __shfl_xor_sync(0xffu, v, 4);
    v += __shfl_xor_sync(0xffu, v, 2);
    v += __shfl_xor_sync(0xffu, v, 1);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}

struct AdamWArgs {
  float lr, beta1, beta2, eps, weight_decay, bias_corr1, bias_corr2_sqrt, max_grad_norm, grad_scale;
};

__device__ __forceinline__ void adamw_one(float& p, float& g, float& m, float& v, const AdamWArgs& a, float coef, float step, float decay,
                                          int zero_grad) {
  const float gr = g * coef;
  m = a.beta1 * m + (1.f - a.beta1) * gr;
  v = a.beta2 * v + (1.f - a.beta2) * gr * gr;
  const float denom = sqrtf(v) / a.bias_corr2_sqrt + a.eps;
  p = p * decay - step * (m / denom);
  if (zero_grad) g = 0.f;
}

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, bf16* __restrict__ p_bf16, int64_t n,
                                                    const float* __restrict__ grad_sumsq, AdamWArgs a, int zero_grad) {
  float coef = a.grad_scale;
  if (grad_sumsq != nullptr && a.max_grad_norm > 0.f) {
    // torch.nn.utils.clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1
    const float total = sqrtf(*grad_sumsq) * a.grad_scale;
    coef *= fminf(1.f, a.max_grad_norm / (total + 1e-6f));
  }
  const float step = a.lr / a.bias_corr1;
  const float decay = 1.f - a.lr * a.weight_decay;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (p_bf16 == nullptr || (reinterpret_cast<uintptr_t>(p_bf16) & 7) == 0);
  const int64_t nvec = vec ? (n >> 2) : 0;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // two float4 of each of the four streams in flight per thread (8 x 16 B loads), then 6-8 x 16 B stores
  for (; i < nvec; i += 2 * stride) {
    const bool two = i + stride < nvec;
    float4 P[2], G[2], M[2], V[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 0 || two) {
        const int64_t j = i + u * stride;
        P[u] = reinterpret_cast<float4*>(p)[j]; G[u] = reinterpret_cast<float4*>(g)[j];
        M[u] = reinterpret_cast<float4*>(m)[j]; V[u] = reinterpret_cast<float4*>(v)[j];
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 0 || two) {
        const int64_t j = i + u * stride;
        adamw_one(P[u].x, G[u].x, M[u].x, V[u].x, a, coef, step, decay, zero_grad);
        adamw_one(P[u].y, G[u].y, M[u].y, V[u].y, a, coef, step, decay, zero_grad);
        adamw_one(P[u].z, G[u].z, M[u].z, V[u].z, a, coef, step, decay, zero_grad);
        adamw_one(P[u].w, G[u].w, M[u].w, V[u].w, a, coef, step, decay, zero_grad);
        reinterpret_cast<float4*>(p)[j] = P[u]; reinterpret_cast<float4*>(m)[j] = M[u]; reinterpret_cast<float4*>(v)[j] = V[u];
        if (zero_grad) reinterpret_cast<float4*>(g)[j] = G[u];
        if (p_bf16) *reinterpret_cast<uint2*>(p_bf16 + 4 * j) = make_uint2(pack_bf16x2(P[u].x, P[u].y), pack_bf16x2(P[u].z, P[u].w));
      }
    }
  }
  for (int64_t j = (nvec << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    float pi = p[j], gi = g[j], mi = m[j], vi = v[j];
    adamw_one(pi, gi, mi, vi, a, coef, step, decay, zero_grad);
    p[j] = pi; m[j] = mi; v[j] = vi;
    if (zero_grad) g[j] = 0.f;
    if (p_bf16) p_bf16[j] = __float2bfloat16_rn(pi);
  }
}

}  // namespace dwb

using namespace dwb;

// Upper bound on the grid of the optimiser-tail kernels (0 = fill the machine).  When the tail runs on a side stream underneath
// the next step's encoder forward (kd.PipelinedTrainer), a machine-filling grid of short CTAs takes register file and thread
// slots away from the persistent one-CTA-per-SM GEMM kernels at every one of their launches and stretches them; a few dozen
// long-running CTAs (256 threads, ~32 registers: co-resident with a GEMM CTA) stream the same bytes under the 75 ms of encoder
// work without ever holding an SM back.
static int g_tail_grid = 0;
extern "C" int dwb_set_tail_grid(int ctas) {
  DWB_CHECK_ARG(ctas >= 0, "dwb_set_tail_grid: negative grid");
  g_tail_grid = ctas;
  return DWB_OK;
}

extern "C" int dwb_grad_sumsq(const float* g, int64_t n, float* out_accum, void* stream) {
  DWB_CHECK_ARG(g && out_accum && n > 0, "dwb_grad_sumsq: bad args");
  DWB_CHECK_ARG((reinterpret_cast<uintptr_t>(g) & 15) == 0, "dwb_grad_sumsq: buffer must be 16 B aligned");
  int64_t blocks = ceil_div64(n / 4 + 1, 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  if (g_tail_grid > 0 && blocks > g_tail_grid) blocks = g_tail_grid;
  sumsq_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(g, n, out_accum);
  DWB_LAUNCH_OK();
  return DWB_OK;
}

extern "C" int dwb_adamw_step(float* p, float* g, float* m, float* v, void* p_bf16, int64_t n, float lr, float beta1, float beta2,
                              float eps, float weight_decay, int step, const float* grad_sumsq, float max_grad_norm,
                              float grad_scale, int zero_grad, void* stream) {
  DWB_CHECK_ARG(p && g && m && v && n > 0 && step >= 1, "dwb_adamw_step: bad args");
  AdamWArgs a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  a.bias_corr1 = 1.f - powf(beta1, (float)step);
  a.bias_corr2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  a.max_grad_norm = max_grad_norm;
  a.grad_scale = grad_scale;
  int64_t blocks = ceil_div64(n, 256 * 8);
  if (blocks < 1) blocks = 1;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  if (g_tail_grid > 0 && blocks > g_tail_grid) blocks = g_tail_grid;
  adamw_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (bf16*)p_bf16, n, grad_sumsq, a, zero_grad);
  DWB_LAUNCH_OK();
  return DWB_OK;
}
