// Library-level entry points: ABI version, status strings, launch counter.
#include <atomic>
#include "common.cuh"

namespace ppy {
thread_local int g_last_cuda_error = 0;
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace ppy

extern "C" {

int ppy_abi_version(void) { return 10; }

const char* ppy_status_string(int status) {
  switch (status) {
    case PPY_OK: return "ok";
    case PPY_ERR_INVALID: return "invalid argument";
    case PPY_ERR_WORKSPACE: return "workspace too small";
    case PPY_ERR_CUDA: return "CUDA call failed";
    case PPY_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown status";
  }
}

int ppy_last_cuda_error(void) { return ppy::g_last_cuda_error; }
long long ppy_kernel_launch_count(void) { return ppy::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
