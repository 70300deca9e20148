// Shared helpers for the PP-YOLO B200 kernel library.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "ppyolo_b200.h"

namespace ppy {

extern thread_local int g_last_cuda_error;
void count_launch(int n = 1);

inline int check_cuda(cudaError_t e) {
  if (e != cudaSuccess) { g_last_cuda_error = (int)e; return PPY_ERR_CUDA; }
  return PPY_OK;
}
inline int check_launch() { count_launch(); return check_cuda(cudaGetLastError()); }

// cudaFuncSetAttribute is per device: `static DeviceOnce once; if (once.first()) { set attribute }` runs once for every device this
// process launches on (a bool would only cover the first one).
struct DeviceOnce {
  unsigned long long seen[2] = {0ull, 0ull};       // up to 128 devices; benign race: at worst the attribute is set twice
  bool first() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 128) return true;
    const unsigned long long bit = 1ull << (dev & 63);
    if (seen[dev >> 6] & bit) return false;
    seen[dev >> 6] |= bit;
    return true;
  }
};

inline cudaStream_t as_stream(ppy_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
inline long long ceil_div(long long a, long long b) { return (a + b - 1) / b; }
inline int dtype_size(int dt) { return dt == PPY_BF16 ? 2 : 4; }

// dtype-generic scalar load/store as float
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == PPY_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == PPY_ACT_LEAKY) return v > 0.f ? v : 0.1f * v;
  if (act == PPY_ACT_MISH) { float sp = (v > 20.f) ? v : log1pf(expf(v)); return v * tanhf(sp); }
  return v;
}

}  // namespace ppy

#define PPY_REQUIRE(cond) do { if (!(cond)) return PPY_ERR_INVALID; } while (0)
