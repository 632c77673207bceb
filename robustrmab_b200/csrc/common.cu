// Error plumbing, launch counter and ABI version of librmab_b200.
#include <stdarg.h>
#include "common.cuh"

namespace rmab {
thread_local char g_err[256] = "";
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(e == cudaErrorNoKernelImageForDevice ? RMAB_E_ARCH : RMAB_E_LAUNCH, "%s: %s", what,
                cudaGetErrorString(e));
  }
  return RMAB_OK;
}
}  // namespace rmab

extern "C" {
int rmab_abi_version(void) { return RMAB_ABI_VERSION; }
const char* rmab_last_error(void) { return rmab::g_err; }
uint64_t rmab_launch_count(void) { return rmab::g_launches.load(); }
}
