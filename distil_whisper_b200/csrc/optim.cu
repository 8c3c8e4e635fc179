// Optimiser tail of the KD step over flat fp32 buffers: gradient sum of squares (for clip_grad_norm_),
// AdamW with the clip coefficient and bias corrections applied on the fly, and the bf16 shadow re-cast that
// the next forward's GEMMs read -- one pass over p/g/m/v.
// Restates ref:training/run_distillation.py:1610-1614 (clip_grad_norm_(max_grad_norm); optimizer.step();
// zero_grad) with torch.optim.AdamW semantics (decoupled weight decay, ref :1402-1407).
#include "common.cuh"

namespace dwb {

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ out) {
  float acc = 0.f;
  const int64_t nvec = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(g)[i];
    acc += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
  }
  for (int64_t i = (nvec << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc += g[i] * g[i];
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
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gr = g[i] * coef;
    const float mi = a.beta1 * m[i] + (1.f - a.beta1) * gr;
    const float vi = a.beta2 * v[i] + (1.f - a.beta2) * gr * gr;
    const float denom = sqrtf(vi) / a.bias_corr2_sqrt + a.eps;
    const float pi = p[i] * decay - step * (mi / denom);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (p_bf16) p_bf16[i] = __float2bfloat16_rn(pi);
    if (zero_grad) g[i] = 0.f;
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
  int64_t blocks = ceil_div64(n, 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  if (g_tail_grid > 0 && blocks > g_tail_grid) blocks = g_tail_grid;
  adamw_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (bf16*)p_bf16, n, grad_sumsq, a, zero_grad);
  DWB_LAUNCH_OK();
  return DWB_OK;
}
