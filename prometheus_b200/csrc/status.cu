// status.cu — ABI version, status strings and the per-thread CUDA error record.
#include <string.h>

#include "common.cuh"

namespace pm {
static thread_local char g_last_error[256] = "";
void set_last_cuda_error(cudaError_t e) {
  const char *s = cudaGetErrorString(e);
  strncpy(g_last_error, s ? s : "unknown", sizeof(g_last_error) - 1);
  g_last_error[sizeof(g_last_error) - 1] = 0;
}
}  // namespace pm

extern "C" {

int pm_abi_version(void) { return PM_ABI_VERSION; }

const char *pm_status_string(int status) {
  switch (status) {
    case PM_OK: return "ok";
    case PM_ERR_BAD_ARG: return "bad argument (null pointer, negative size or misaligned buffer)";
    case PM_ERR_UNSUPPORTED: return "unsupported configuration";
    case PM_ERR_WORKSPACE: return "workspace too small";
    case PM_ERR_CUDA: return "CUDA runtime error";
    case PM_ERR_RANGE: return "size exceeds an implementation limit";
    default: return "unknown status";
  }
}

const char *pm_last_cuda_error(void) { return pm::g_last_error; }

}  // extern "C"
