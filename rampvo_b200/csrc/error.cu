// error.cu — error reporting and device probing for librampvo_b200.
#include <stdarg.h>

#include "common.cuh"

namespace rvo {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  cudaGetLastError();  // clear the sticky-less error so the next call starts clean
  return RVO_ERR_CUDA;
}

}  // namespace rvo

extern "C" int rvo_abi_version(void) { return RVO_ABI_VERSION; }

extern "C" const char* rvo_last_error(void) { return rvo::g_err; }

extern "C" uint64_t rvo_launch_count(void) { return rvo::g_launches; }

namespace rvo {
thread_local int g_sm_budget = kNumSMs;
}

extern "C" int rvo_set_sm_budget(int n_sms) {
  RVO_CHECK_ARG(n_sms >= 0 && n_sms <= rvo::kNumSMs, "rvo_set_sm_budget: %d (1..%d, 0 = all)", n_sms, rvo::kNumSMs);
  rvo::g_sm_budget = n_sms == 0 ? rvo::kNumSMs : n_sms;
  return RVO_OK;
}

extern "C" int rvo_get_sm_budget(void) { return rvo::g_sm_budget; }

extern "C" int rvo_device_cc(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return -1; }
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return major * 10 + minor;
}
