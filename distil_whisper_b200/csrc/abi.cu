// C-ABI plumbing: error text, version, device probe.  See include/dwb.h for the contract.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

extern "C" void dwb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* dwb_last_error(void) { return g_err; }
extern "C" int dwb_abi_version(void) { return 2; }

// kernel launches issued by this library since the last reset (host-side count of <<<>>> launches; a CUDA-graph replay
// re-launches the captured kernels without passing through here)
static unsigned long long g_launches = 0;
extern "C" void dwb_count_launch(void) { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
extern "C" int64_t dwb_launch_count(int reset) {
  const unsigned long long v = reset ? __atomic_exchange_n(&g_launches, 0ull, __ATOMIC_RELAXED) : __atomic_load_n(&g_launches, __ATOMIC_RELAXED);
  return (int64_t)v;
}

// 0 when a compute-capability 10.x device is current; error otherwise (the product never falls back to a CPU path)
extern "C" int dwb_check_device(void) {
  int dev = 0;
  DWB_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  DWB_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    dwb_set_error("distil-whisper-b200 kernels are built for sm_100a only; current device is sm_%d%d (%s)", prop.major, prop.minor,
                  prop.name);
    return DWB_ERR_UNSUPPORTED;
  }
  return DWB_OK;
}

// Direction in which the next GEMM / LayerNorm / attention-forward launches walk the rows of their [rows, features] operands:
// 0 = first to last (default), 1 = last to first.  A chain of kernels that stream a tensor larger than the 126 MB L2 alternates the
// direction, so that every consumer starts on the rows its producer wrote last (still in L2) instead of on the ones already evicted
// (engine.encoder_forward).  Host-side launch parameter; not thread-safe across concurrently launching host threads.
static int g_row_walk_reverse = 0;
extern "C" int dwb_row_walk_reverse(void) { return g_row_walk_reverse; }
extern "C" int dwb_set_row_walk(int reverse) {
  g_row_walk_reverse = reverse ? 1 : 0;
  return DWB_OK;
}
