// Glue kernels of the fp32-grade tensor-core path (PPY_F16X2 activations: an fp32-grade value carried as an fp16 hi/lo pair in
// two planes, see ppyolo_b200.h): MaxPool 3x3/s2, the vd shortcut's AvgPool 2x2, SPP.  HBM-bound like their bf16 siblings in
// layout_pool.cu; one thread moves one 8-channel vector = a 16-byte load/store per plane.  hi + lo is exact in fp32 (the pair
// was split from an fp32 number), so max-type ops compare and re-split exact values -- results are value-identical to the
// fp32 kernels on the joined tensors; the average adds in the reference's operation order and re-splits its fp32 result.
#include <cuda_fp16.h>
#include <float.h>
#include "common.cuh"

namespace ppy {
namespace {

__device__ __forceinline__ void load_pair8(const __half* p, long long plane, float (&v)[8]) {
  const uint4 hi = __ldg(reinterpret_cast<const uint4*>(p));
  const uint4 lo = __ldg(reinterpret_cast<const uint4*>(p + plane));
  const __half2* h = reinterpret_cast<const __half2*>(&hi);
  const __half2* l = reinterpret_cast<const __half2*>(&lo);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 a = __half22float2(h[i]), b = __half22float2(l[i]);
    v[2 * i] = a.x + b.x;
    v[2 * i + 1] = a.y + b.y;
  }
}

__device__ __forceinline__ void store_pair8(__half* p, long long plane, const float (&v)[8]) {
  uint4 hi, lo;
  __half2* h = reinterpret_cast<__half2*>(&hi);
  __half2* l = reinterpret_cast<__half2*>(&lo);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    const float2 hf = __half22float2(h[i]);
    l[i] = __floats2half2_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
  }
  *reinterpret_cast<uint4*>(p) = hi;
  *reinterpret_cast<uint4*>(p + plane) = lo;
}

inline unsigned grid_for(long long work, int threads) {
  long long b = ceil_div(work, threads);
  const long long cap = 148ll * 32;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

// MODE 0 = max 3x3 s2 p1 (model/resnet_vd.py:103), 1 = avg 2x2 s2 p0 (model/resnet_vd.py:30)
template <int MODE>
__global__ void __launch_bounds__(256) pool_pair_kernel(const __half* __restrict__ x, int x_ld, long long x_plane, __half* __restrict__ y,
                                                        int y_ld, long long y_plane, int n, int h, int w, int c, int ho, int wo) {
  const int cv = c >> 3;
  const unsigned total = (unsigned)(n * ho * wo * cv);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = (int)(i % (unsigned)cv);
    unsigned p = i / (unsigned)cv;
    const int ox = (int)(p % (unsigned)wo); p /= (unsigned)wo;
    const int oy = (int)(p % (unsigned)ho);
    const int img = (int)(p / (unsigned)ho);
    float acc[8];
    if (MODE == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = -FLT_MAX;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const int iy = oy * 2 + dy;
        if (iy < 0 || iy >= h) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int ix = ox * 2 + dx;
          if (ix < 0 || ix >= w) continue;
          float t[8];
          load_pair8(x + (((long long)img * h + iy) * w + ix) * x_ld + v * 8, x_plane, t);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] = fmaxf(acc[k], t[k]);
        }
      }
    } else {
      float a[8], b[8], cc[8], d[8];
      const __half* base = x + (((long long)img * h + oy * 2) * w + ox * 2) * x_ld + v * 8;
      load_pair8(base, x_plane, a);
      load_pair8(base + x_ld, x_plane, b);
      load_pair8(base + (long long)w * x_ld, x_plane, cc);
      load_pair8(base + (long long)w * x_ld + x_ld, x_plane, d);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(a[k], b[k]), cc[k]), d[k]), 4.f);
    }
    store_pair8(y + (((long long)img * ho + oy) * wo + ox) * y_ld + v * 8, y_plane, acc);
  }
}

// MaxPool 3x3 / stride 2 / pad 1, second version: one thread = one 8-channel vector of TWO vertically adjacent output pixels.
// The five input rows they share are read once (15 positions x 2 planes instead of 18 per output), coordinates are clamped to
// the image instead of tested (a clamped position always lies inside the window, so the maximum is unchanged): branch-free, all
// 30 loads of a thread in flight together.
__global__ void __launch_bounds__(256) maxpool_pair2_kernel(const __half* __restrict__ x, int x_ld, long long x_plane, __half* __restrict__ y,
                                                            int y_ld, long long y_plane, int n, int h, int w, int c, int ho, int wo) {
  const int cv = c >> 3, ho2 = (ho + 1) >> 1;
  const unsigned total = (unsigned)(n * ho2 * wo * cv);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = (int)(i % (unsigned)cv);
    unsigned p = i / (unsigned)cv;
    const int ox = (int)(p % (unsigned)wo); p /= (unsigned)wo;
    const int oy = 2 * (int)(p % (unsigned)ho2);
    const int img = (int)(p / (unsigned)ho2);
    const __half* base = x + (long long)img * h * w * x_ld + v * 8;
    int xs[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) { const int t = ox * 2 + d - 1; xs[d] = t < 0 ? 0 : (t >= w ? w - 1 : t); }
    float rmax[5][8];                      // horizontal 3-max of input rows 2*oy - 1 .. 2*oy + 3
#pragma unroll
    for (int r = 0; r < 5; ++r) {
      int yy = oy * 2 + r - 1;
      yy = yy < 0 ? 0 : (yy >= h ? h - 1 : yy);
      float a[8], b[8], cc[8];
      const __half* row = base + (long long)yy * w * x_ld;
      load_pair8(row + (long long)xs[0] * x_ld, x_plane, a);
      load_pair8(row + (long long)xs[1] * x_ld, x_plane, b);
      load_pair8(row + (long long)xs[2] * x_ld, x_plane, cc);
#pragma unroll
      for (int k = 0; k < 8; ++k) rmax[r][k] = fmaxf(fmaxf(a[k], b[k]), cc[k]);
    }
    float o0[8], o1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      // output row oy: input rows 2oy-1..2oy+1 = r 0..2; output row oy+1: input rows 2oy+1..2oy+3 = r 2..4.  Rows clamped at the
      // bottom repeat row h-1, which is inside the window of output row oy+1 whenever that row exists.
      o0[k] = fmaxf(fmaxf(rmax[0][k], rmax[1][k]), rmax[2][k]);
      o1[k] = fmaxf(fmaxf(rmax[2][k], rmax[3][k]), rmax[4][k]);
    }
    __half* dst = y + (((long long)img * ho + oy) * wo + ox) * y_ld + v * 8;
    store_pair8(dst, y_plane, o0);
    if (oy + 1 < ho) store_pair8(dst + (long long)wo * y_ld, y_plane, o1);
  }
}

// SPP (model/custom_layers.py:275-290), separable + cascaded like spp_separable_kernel: maxpool9 = maxpool5(maxpool5(x)),
// maxpool13 = maxpool5(maxpool9), each 5x5 as a row pass and a column pass over fp32 values in shared memory.
// One CTA = one image x SPPP_VECS 8-channel vectors.
constexpr int SPPP_VECS = 2;

__device__ __forceinline__ void vmax8(float (&m)[8], const float4* s) {
  const float4 a = s[0], b = s[1];
  m[0] = fmaxf(m[0], a.x); m[1] = fmaxf(m[1], a.y); m[2] = fmaxf(m[2], a.z); m[3] = fmaxf(m[3], a.w);
  m[4] = fmaxf(m[4], b.x); m[5] = fmaxf(m[5], b.y); m[6] = fmaxf(m[6], b.z); m[7] = fmaxf(m[7], b.w);
}
__device__ __forceinline__ void put8(float4* s, const float (&m)[8]) {
  s[0] = make_float4(m[0], m[1], m[2], m[3]);
  s[1] = make_float4(m[4], m[5], m[6], m[7]);
}
__device__ __forceinline__ void get8(const float4* s, float (&m)[8]) {
  const float4 a = s[0], b = s[1];
  m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w; m[4] = b.x; m[5] = b.y; m[6] = b.z; m[7] = b.w;
}

__global__ void __launch_bounds__(256) spp_pair_kernel(const __half* __restrict__ x, int x_ld, long long x_plane, __half* __restrict__ y,
                                                       int y_ld, long long y_plane, int h, int w, int c) {
  extern __shared__ float4 sppp_smem[];
  const int hw = h * w;
  float4* cur = sppp_smem;                       // [hw][SPPP_VECS][2 float4]
  float4* tmp = sppp_smem + hw * SPPP_VECS * 2;
  const int img = blockIdx.y;
  const int c0 = blockIdx.x * SPPP_VECS * 8;
  const __half* xi = x + (long long)img * hw * x_ld + c0;
  __half* yo = y + (long long)img * hw * y_ld + c0;
  const int items = hw * SPPP_VECS;
  for (int i = threadIdx.x; i < items; i += blockDim.x) {
    const int pix = i / SPPP_VECS, v = i % SPPP_VECS;
    float t[8];
    load_pair8(xi + (long long)pix * x_ld + v * 8, x_plane, t);
    put8(cur + 2 * i, t);
    store_pair8(yo + (long long)pix * y_ld + v * 8, y_plane, t);
  }
  __syncthreads();
  for (int round = 1; round <= 3; ++round) {
    for (int i = threadIdx.x; i < items; i += blockDim.x) {      // row pass
      const int pix = i / SPPP_VECS, v = i % SPPP_VECS;
      const int px = pix % w, row0 = pix - px;
      float m[8];
      get8(cur + 2 * i, m);
      for (int d = -2; d <= 2; ++d) {
        const int xx = px + d;
        if (d != 0 && xx >= 0 && xx < w) vmax8(m, cur + 2 * ((row0 + xx) * SPPP_VECS + v));
      }
      put8(tmp + 2 * i, m);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < items; i += blockDim.x) {      // column pass
      const int pix = i / SPPP_VECS, v = i % SPPP_VECS;
      const int py = pix / w;
      float m[8];
      get8(tmp + 2 * i, m);
      for (int d = -2; d <= 2; ++d) {
        const int yy = py + d;
        if (d != 0 && yy >= 0 && yy < h) vmax8(m, tmp + 2 * ((pix + d * w) * SPPP_VECS + v));
      }
      put8(cur + 2 * i, m);
      store_pair8(yo + (long long)pix * y_ld + round * c + v * 8, y_plane, m);
    }
    __syncthreads();
  }
}

// NHWC fp32 -> pair planes / pair planes -> NHWC fp32 (module-level interop and tests; rows = n*h*w pixels)
__global__ void split_rows_kernel(const float* __restrict__ x, int x_ld, __half* __restrict__ y, int y_ld, long long y_plane,
                                  long long rows, int c) {
  const long long total = rows * c;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c;
    const int ch = (int)(i % c);
    const float f = __ldg(x + r * x_ld + ch);
    const __half hi = __float2half_rn(f);
    y[r * y_ld + ch] = hi;
    y[y_plane + r * y_ld + ch] = __float2half_rn(f - __half2float(hi));
  }
}
__global__ void join_rows_kernel(const __half* __restrict__ x, int x_ld, long long x_plane, float* __restrict__ y, int y_ld,
                                 long long rows, int c) {
  const long long total = rows * c;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c;
    const int ch = (int)(i % c);
    y[r * y_ld + ch] = __half2float(x[r * x_ld + ch]) + __half2float(x[x_plane + r * x_ld + ch]);
  }
}

inline bool pair_ok(const void* p, int ld, long long plane, int c) {
  return p && (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % 8 == 0 && c % 8 == 0 && c > 0 && ld >= c && plane > 0 && plane % 8 == 0;
}

}  // namespace
}  // namespace ppy

extern "C" {
using namespace ppy;

int ppy_maxpool3x3s2_f16x2(const void* x, int x_ld, long long x_plane, void* y, int y_ld, long long y_plane, int n, int h, int w,
                           int c, ppy_stream_t s) {
  PPY_REQUIRE(n > 0 && h > 0 && w > 0 && pair_ok(x, x_ld, x_plane, c) && pair_ok(y, y_ld, y_plane, c));
  const int ho = (h + 1) / 2, wo = (w + 1) / 2;
  const long long total = (long long)n * ho * wo * (c / 8);
  PPY_REQUIRE(total < 0x7FFFFFFFll);
  const long long total2 = (long long)n * ((ho + 1) / 2) * wo * (c / 8);
  maxpool_pair2_kernel<<<grid_for(total2, 256), 256, 0, as_stream(s)>>>((const __half*)x, x_ld, x_plane, (__half*)y, y_ld, y_plane, n, h, w,
                                                                         c, ho, wo);
  return check_launch();
}

int ppy_avgpool2x2_f16x2(const void* x, int x_ld, long long x_plane, void* y, int y_ld, long long y_plane, int n, int h, int w, int c,
                         ppy_stream_t s) {
  PPY_REQUIRE(n > 0 && h > 1 && w > 1 && pair_ok(x, x_ld, x_plane, c) && pair_ok(y, y_ld, y_plane, c));
  const int ho = h / 2, wo = w / 2;
  const long long total = (long long)n * ho * wo * (c / 8);
  PPY_REQUIRE(total < 0x7FFFFFFFll);
  pool_pair_kernel<1><<<grid_for(total, 256), 256, 0, as_stream(s)>>>((const __half*)x, x_ld, x_plane, (__half*)y, y_ld, y_plane, n, h, w,
                                                                       c, ho, wo);
  return check_launch();
}

int ppy_spp_f16x2(const void* x, int x_ld, long long x_plane, void* y, int y_ld, long long y_plane, int n, int h, int w, int c,
                  ppy_stream_t s) {
  PPY_REQUIRE(n > 0 && h > 0 && w > 0 && pair_ok(x, x_ld, x_plane, c) && pair_ok(y, y_ld, y_plane, c) && y_ld >= 4 * c);
  const size_t smem = (size_t)2 * h * w * SPPP_VECS * 32;
  PPY_REQUIRE(c % (SPPP_VECS * 8) == 0 && smem <= 200 * 1024);
  if (smem > 48 * 1024) {
    if (check_cuda(cudaFuncSetAttribute(spp_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return PPY_ERR_CUDA;
  }
  dim3 grid((unsigned)(c / (SPPP_VECS * 8)), (unsigned)n);
  spp_pair_kernel<<<grid, 256, smem, as_stream(s)>>>((const __half*)x, x_ld, x_plane, (__half*)y, y_ld, y_plane, h, w, c);
  return check_launch();
}

int ppy_split_f16x2(const float* x, int x_ld, void* y, int y_ld, long long y_plane, long long rows, int c, ppy_stream_t s) {
  PPY_REQUIRE(x && y && rows > 0 && c > 0 && x_ld >= c && y_ld >= c && y_plane > 0);
  split_rows_kernel<<<grid_for(rows * c, 256), 256, 0, as_stream(s)>>>(x, x_ld, (__half*)y, y_ld, y_plane, rows, c);
  return check_launch();
}

int ppy_join_f16x2(const void* x, int x_ld, long long x_plane, float* y, int y_ld, long long rows, int c, ppy_stream_t s) {
  PPY_REQUIRE(x && y && rows > 0 && c > 0 && x_ld >= c && y_ld >= c && x_plane > 0);
  join_rows_kernel<<<grid_for(rows * c, 256), 256, 0, as_stream(s)>>>((const __half*)x, x_ld, x_plane, y, y_ld, rows, c);
  return check_launch();
}

}  // extern "C"
