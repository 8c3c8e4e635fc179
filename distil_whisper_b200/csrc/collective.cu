// The KD step's only collective -- the sum all-reduce of the flat fp32 student gradient (ref:training/run_distillation.py:1609,
// DDP's implicit all-reduce inside accelerator.backward) -- as ONE small-footprint kernel over NVLink / NVSwitch peer memory, so that it
// can run on a side stream underneath the next step's encoder forward without taking SMs away from the persistent GEMM kernels
// (NCCL's channel CTAs cannot co-reside with a 200 KB-smem GEMM CTA; a 512-thread, no-smem CTA can).
//
// The gradient buffer lives in symmetric memory (every rank maps every peer's buffer; torch.distributed._symmetric_memory does the
// allocation, the handle exchange and the device-side barriers before and after this kernel -- plumbing).  Rank r owns slice r of the
// vector ("two-shot", in place):
//   NVSwitch multicast available : v = multimem.ld_reduce.add(slice r)  -- the switch sums the N copies in flight --, then
//                                  multimem.st(slice r, v) writes the sum into all N buffers: every byte crosses the fabric once each way.
//   otherwise (P2P only)         : v = sum over peers of a plain load of their slice r, then a plain store into every peer's slice r.
// No rank reads a slice it does not own, so the update is race-free between the two barriers.
#include "common.cuh"

namespace dwb {

constexpr int AR_THREADS = 512;
constexpr int AR_MAX_RANKS = 16;

struct PeerPtrs {
  float* p[AR_MAX_RANKS];
};

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// vec4 index range [lo, hi) of this rank's slice
__global__ void __launch_bounds__(AR_THREADS) allreduce_nvls_kernel(float* __restrict__ mc, int64_t lo, int64_t hi) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < hi; i += 4 * stride) {          // four independent 16 B fabric reads in flight per thread
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = multimem_ld_reduce_add(mc + 4 * (i + u * stride));
#pragma unroll
    for (int u = 0; u < 4; ++u) multimem_st(mc + 4 * (i + u * stride), v[u]);
  }
  for (; i < hi; i += stride) multimem_st(mc + 4 * i, multimem_ld_reduce_add(mc + 4 * i));
}

__global__ void __launch_bounds__(AR_THREADS) allreduce_p2p_kernel(PeerPtrs peers, int world, int64_t lo, int64_t hi) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int r = 0; r < world; ++r) {                     // fixed rank order: every rank's slice is summed identically
      const float4 v = reinterpret_cast<const float4*>(peers.p[r])[i];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    for (int r = 0; r < world; ++r) reinterpret_cast<float4*>(peers.p[r])[i] = acc;
  }
}

}  // namespace dwb

using namespace dwb;

// n: elements (multiple of 4, buffers 16 B aligned).  multicast_ptr: the NVSwitch multicast address of the buffer, or NULL ->
// peer_ptrs[world] (every rank's mapping of every buffer, own rank included) are used with plain loads / stores.  The caller
// brackets the call with device-side barriers across the ranks (all gradients complete before; all slices written after).
extern "C" int dwb_allreduce_symm(void* multicast_ptr, const void* const* peer_ptrs, int rank, int world, int64_t n, int max_ctas, void* stream) {
  DWB_CHECK_ARG(world >= 1 && world <= AR_MAX_RANKS && rank >= 0 && rank < world, "dwb_allreduce_symm: bad rank %d / world %d", rank, world);
  DWB_CHECK_ARG(n > 0 && (n % 4) == 0, "dwb_allreduce_symm: n=%lld must be a positive multiple of 4", (long long)n);
  DWB_CHECK_ARG(multicast_ptr != nullptr || peer_ptrs != nullptr, "dwb_allreduce_symm: need a multicast pointer or the peer pointers");
  const int64_t nvec = n / 4;
  const int64_t per = ceil_div64(nvec, world);
  const int64_t lo = per * rank, hi = lo + per < nvec ? lo + per : nvec;
  if (hi <= lo) return DWB_OK;
  int ctas = max_ctas > 0 ? max_ctas : 32;
  const int64_t need = ceil_div64(hi - lo, AR_THREADS);
  if (need < ctas) ctas = (int)need;
  if (multicast_ptr != nullptr) {
    DWB_CHECK_ARG((reinterpret_cast<uintptr_t>(multicast_ptr) & 15) == 0, "dwb_allreduce_symm: multicast pointer not 16 B aligned");
    allreduce_nvls_kernel<<<ctas, AR_THREADS, 0, (cudaStream_t)stream>>>(reinterpret_cast<float*>(multicast_ptr), lo, hi);
  } else {
    PeerPtrs pp;
    for (int r = 0; r < world; ++r) {
      DWB_CHECK_ARG(peer_ptrs[r] != nullptr && (reinterpret_cast<uintptr_t>(peer_ptrs[r]) & 15) == 0, "dwb_allreduce_symm: peer pointer %d null / unaligned", r);
      pp.p[r] = reinterpret_cast<float*>(const_cast<void*>(peer_ptrs[r]));
    }
    allreduce_p2p_kernel<<<ctas, AR_THREADS, 0, (cudaStream_t)stream>>>(pp, world, lo, hi);
  }
  DWB_LAUNCH_OK();
  return DWB_OK;
}
