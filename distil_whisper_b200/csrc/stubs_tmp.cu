// TEMPORARY: entry points whose kernels land in the next commits.
#include "common.cuh"
extern "C" int dwb_attention_fwd(const void*, int64_t, const void*, int64_t, const void*, int64_t, void*, int64_t, float*, int, int, int, int, int, int, float, void*);
extern "C" int dwb_attention_fwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo,
                                    float* lse, int B, int H, int Sq, int Sk, int head_dim, int causal, float scale, void* stream) {
  dwb_set_error("dwb_attention_fwd_tc: not built yet");
  return DWB_ERR_UNSUPPORTED;
}
