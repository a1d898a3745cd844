// Library plumbing: status strings, CUDA error capture, device properties.
#include <atomic>
#include <cstdio>
#include <cstring>

#include "vu_common.cuh"

namespace vu {

static thread_local char g_last_error[256] = "";
static std::atomic<unsigned long long> g_launches{0};

void note_launch(int kernels) { g_launches.fetch_add((unsigned long long)kernels, std::memory_order_relaxed); }

int record_cuda(cudaError_t e) {
  if (e == cudaSuccess) return VU_OK;
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
  return VU_ERR_CUDA;
}

int device_sms() {
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev] = v;
  }
  return sms[dev];
}

}  // namespace vu

extern "C" int vu_abi_version(void) { return VU_ABI_VERSION; }

extern "C" const char* vu_status_string(int status) {
  switch (status) {
    case VU_OK: return "ok";
    case VU_ERR_INVALID_ARG: return "invalid argument";
    case VU_ERR_UNSUPPORTED: return "unsupported size, alignment or parameter";
    case VU_ERR_WORKSPACE: return "workspace too small or misaligned";
    case VU_ERR_CUDA: return "CUDA error (see vu_last_cuda_error)";
    default: return "unknown status";
  }
}

extern "C" const char* vu_last_cuda_error(void) { return vu::g_last_error; }

extern "C" uint64_t vu_launch_count(void) { return vu::g_launches.load(); }
