// HBM-bound glue kernels on NHWC activations: layout conversion, pooling, SPP, nearest upsample,
// channel-slice copies (concat), CoordConv channels.  All move 16-byte vectors along the channel
// dimension (8 bf16 / 4 fp32) so every warp access is a run of full 32-byte sectors; grids are
// grid-stride loops capped at a few waves of the 148 SMs.
#include <float.h>
#include <math_constants.h>
#include "common.cuh"

namespace ppy {
namespace {

template <typename T> struct Vec16 { static constexpr int N = 16 / sizeof(T); };

template <typename T>
__device__ __forceinline__ void load_vec(const T* p, float (&v)[Vec16<T>::N]) {
  uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
  const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
  for (int i = 0; i < Vec16<T>::N; ++i) v[i] = to_f<T>(e[i]);
}
template <typename T>
__device__ __forceinline__ void store_vec(T* p, const float (&v)[Vec16<T>::N]) {
  uint4 raw;
  T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
  for (int i = 0; i < Vec16<T>::N; ++i) e[i] = from_f<T>(v[i]);
  *reinterpret_cast<uint4*>(p) = raw;
}

inline unsigned grid_for(long long work, int threads) {
  long long b = ceil_div(work, threads);
  long long cap = 148ll * 32;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, T* __restrict__ y, int n, int c, int h, int w, int ld) {
  const long long pixels = (long long)n * h * w;
  const long long hw = (long long)h * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += (long long)gridDim.x * blockDim.x) {
    const long long img = i / hw, rem = i % hw;
    const float* src = x + img * c * hw + rem;
    T* dst = y + i * ld;
    for (int ch = 0; ch < ld; ++ch) dst[ch] = from_f<T>(ch < c ? __ldg(src + ch * hw) : 0.f);
  }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ x, int ld, float* __restrict__ y, int n, int c, int h, int w) {
  const long long total = (long long)n * c * h * w;
  const long long hw = (long long)h * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long img = i / (c * hw);
    const int ch = (int)((i / hw) % c);
    const long long rem = i % hw;
    y[i] = to_f<T>(x[(img * hw + rem) * ld + ch]);
  }
}

// generic window op: MODE 0 = max 3x3 s2 p1, 1 = avg 2x2 s2 p0
template <typename T, int MODE>
__global__ void pool_kernel(const T* __restrict__ x, int x_ld, T* __restrict__ y, int y_ld, int n, int h, int w, int c,
                            int ho, int wo) {
  constexpr int V = Vec16<T>::N;
  const int cv = c / V;
  const unsigned total = (unsigned)(n * ho * wo * cv);          // host guarantees < 2^31: 32-bit index math
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = (int)(i % (unsigned)cv);
    unsigned p = i / (unsigned)cv;
    const int ox = (int)(p % (unsigned)wo); p /= (unsigned)wo;
    const int oy = (int)(p % (unsigned)ho);
    const int img = (int)(p / (unsigned)ho);
    float acc[V];
    if (MODE == 0) {
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] = -FLT_MAX;
      for (int dy = -1; dy <= 1; ++dy) {
        const int iy = oy * 2 + dy;
        if (iy < 0 || iy >= h) continue;
        for (int dx = -1; dx <= 1; ++dx) {
          const int ix = ox * 2 + dx;
          if (ix < 0 || ix >= w) continue;
          float t[V];
          load_vec<T>(x + (((long long)img * h + iy) * w + ix) * x_ld + v * V, t);
#pragma unroll
          for (int k = 0; k < V; ++k) acc[k] = fmaxf(acc[k], t[k]);
        }
      }
    } else {
      float a[V], b[V], cc[V], d[V];
      const T* base = x + (((long long)img * h + oy * 2) * w + ox * 2) * x_ld + v * V;
      load_vec<T>(base, a);
      load_vec<T>(base + x_ld, b);
      load_vec<T>(base + (long long)w * x_ld, cc);
      load_vec<T>(base + (long long)w * x_ld + x_ld, d);
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(a[k], b[k]), cc[k]), d[k]), 4.f);
    }
    store_vec<T>(y + (((long long)img * ho + oy) * wo + ox) * y_ld + v * V, acc);
  }
}

// bf16 fast path of both pools: one thread = one 16-byte channel vector of TWO horizontally adjacent output pixels; every
// load of the thread (8 for the average, up to 15 for the max) is issued before the first use, max runs on packed bf16x2
// (exact), the average in fp32 with the reference's operation order.  MODE 0 = max 3x3 s2 p1, 1 = avg 2x2 s2 p0.
__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint4 max16(uint4 a, uint4 b) {
  return make_uint4(max_bf16x2(a.x, b.x), max_bf16x2(a.y, b.y), max_bf16x2(a.z, b.z), max_bf16x2(a.w, b.w));
}
__device__ __forceinline__ uint32_t avg4_bf16x2(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  const float lo = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(__uint_as_float(a << 16), __uint_as_float(b << 16)), __uint_as_float(c << 16)),
                                       __uint_as_float(d << 16)), 4.f);
  const float hi = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(__uint_as_float(a & 0xFFFF0000u), __uint_as_float(b & 0xFFFF0000u)),
                                                 __uint_as_float(c & 0xFFFF0000u)), __uint_as_float(d & 0xFFFF0000u)), 4.f);
  __nv_bfloat162 r = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&r);
}

template <int MODE>
__global__ void __launch_bounds__(256) pool2_bf16_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, __nv_bfloat16* __restrict__ y, int y_ld,
                                                         int n, int h, int w, int c, int ho, int wo) {
  const int cv = c >> 3, wp = (wo + 1) >> 1;                     // channel vectors, output pixel pairs per row
  const unsigned total = (unsigned)(n * ho * wp * cv);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = (int)(i % (unsigned)cv);
    unsigned p = i / (unsigned)cv;
    const int ox = 2 * (int)(p % (unsigned)wp); p /= (unsigned)wp;
    const int oy = (int)(p % (unsigned)ho);
    const int img = (int)(p / (unsigned)ho);
    const bool two = ox + 1 < wo;
    __nv_bfloat16* dst = y + (((long long)img * ho + oy) * wo + ox) * y_ld + v * 8;
    if (MODE == 1) {
      const __nv_bfloat16* base = x + (((long long)img * h + oy * 2) * w + ox * 2) * x_ld + v * 8;
      const __nv_bfloat16* base2 = base + (long long)w * x_ld;
      uint4 a[4], b[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool ok = k < 2 || two;
        a[k] = ok ? ldg16(base + k * x_ld) : make_uint4(0u, 0u, 0u, 0u);
        b[k] = ok ? ldg16(base2 + k * x_ld) : make_uint4(0u, 0u, 0u, 0u);
      }
      *reinterpret_cast<uint4*>(dst) = make_uint4(avg4_bf16x2(a[0].x, a[1].x, b[0].x, b[1].x), avg4_bf16x2(a[0].y, a[1].y, b[0].y, b[1].y),
                                                  avg4_bf16x2(a[0].z, a[1].z, b[0].z, b[1].z), avg4_bf16x2(a[0].w, a[1].w, b[0].w, b[1].w));
      if (two)
        *reinterpret_cast<uint4*>(dst + y_ld) = make_uint4(avg4_bf16x2(a[2].x, a[3].x, b[2].x, b[3].x), avg4_bf16x2(a[2].y, a[3].y, b[2].y, b[3].y),
                                                           avg4_bf16x2(a[2].z, a[3].z, b[2].z, b[3].z), avg4_bf16x2(a[2].w, a[3].w, b[2].w, b[3].w));
    } else {
      // input columns 2ox-1 .. 2ox+3 (5), rows 2oy-1 .. 2oy+1 (3); out-of-image taps are replaced by an in-window valid tap
      uint4 t[3][5];
      const int cx = ox * 2, cy = oy * 2;                        // always-valid centre of the first window
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        int iy = cy - 1 + r;
        iy = (iy < 0 || iy >= h) ? cy : iy;
        const __nv_bfloat16* row = x + ((long long)img * h + iy) * w * x_ld + v * 8;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          int ix = cx - 1 + k;
          // columns 0..2 belong to window 0 (fallback: its centre cx), 2..4 to window 1 (fallback: its centre cx + 2, if it exists)
          if (ix < 0 || ix >= w) ix = (k < 2 || !two) ? cx : cx + 2;
          if (!two && k > 2) ix = cx;
          t[r][k] = ldg16(row + (long long)ix * x_ld);
        }
      }
      uint4 m0 = t[0][0], m1 = t[0][2];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { m0 = max16(m0, t[r][k]); m1 = max16(m1, t[r][k + 2]); }
      }
      *reinterpret_cast<uint4*>(dst) = m0;
      if (two) *reinterpret_cast<uint4*>(dst + y_ld) = m1;
    }
  }
}

// SPP: one thread per (pixel, 16-byte channel vector); max over nested 5/9/13 windows in a single sweep
// of the 13x13 neighbourhood (each tap is classified into the smallest window containing it).
template <typename T>
__global__ void spp_kernel(const T* __restrict__ x, int x_ld, T* __restrict__ y, int y_ld, int n, int h, int w, int c) {
  constexpr int V = Vec16<T>::N;
  const int cv = c / V;
  const long long total = (long long)n * h * w * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    long long p = i / cv;
    const int ox = (int)(p % w); p /= w;
    const int oy = (int)(p % h);
    const int img = (int)(p / h);
    float m5[V], m9[V], m13[V], ctr[V];
#pragma unroll
    for (int k = 0; k < V; ++k) { m5[k] = -FLT_MAX; m9[k] = -FLT_MAX; m13[k] = -FLT_MAX; }
    for (int dy = -6; dy <= 6; ++dy) {
      const int iy = oy + dy;
      if (iy < 0 || iy >= h) continue;
      const int ady = dy < 0 ? -dy : dy;
      for (int dx = -6; dx <= 6; ++dx) {
        const int ix = ox + dx;
        if (ix < 0 || ix >= w) continue;
        const int adx = dx < 0 ? -dx : dx;
        const int r = ady > adx ? ady : adx;
        float t[V];
        load_vec<T>(x + (((long long)img * h + iy) * w + ix) * x_ld + v * V, t);
        if (r <= 2) {
#pragma unroll
          for (int k = 0; k < V; ++k) m5[k] = fmaxf(m5[k], t[k]);
        } else if (r <= 4) {
#pragma unroll
          for (int k = 0; k < V; ++k) m9[k] = fmaxf(m9[k], t[k]);
        } else {
#pragma unroll
          for (int k = 0; k < V; ++k) m13[k] = fmaxf(m13[k], t[k]);
        }
        if (r == 0) {
#pragma unroll
          for (int k = 0; k < V; ++k) ctr[k] = t[k];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) { m9[k] = fmaxf(m9[k], m5[k]); m13[k] = fmaxf(m13[k], m9[k]); }
    T* dst = y + (((long long)img * h + oy) * w + ox) * y_ld + v * V;
    store_vec<T>(dst, ctr);
    store_vec<T>(dst + c, m5);
    store_vec<T>(dst + 2 * c, m9);
    store_vec<T>(dst + 3 * c, m13);
  }
}

// SPP, separable + cascaded: maxpool9 = maxpool5(maxpool5(x)), maxpool13 = maxpool5(maxpool9) (exact for max with
// -inf padding), each 5x5 as a row pass and a column pass in shared memory.  One CTA = one image x 4 channel vectors.
template <typename T> __device__ __forceinline__ uint4 vmax(uint4 a, uint4 b);
template <> __device__ __forceinline__ uint4 vmax<float>(uint4 a, uint4 b) {
  return make_uint4(__float_as_uint(fmaxf(__uint_as_float(a.x), __uint_as_float(b.x))), __float_as_uint(fmaxf(__uint_as_float(a.y), __uint_as_float(b.y))),
                    __float_as_uint(fmaxf(__uint_as_float(a.z), __uint_as_float(b.z))), __float_as_uint(fmaxf(__uint_as_float(a.w), __uint_as_float(b.w))));
}
template <> __device__ __forceinline__ uint4 vmax<__nv_bfloat16>(uint4 a, uint4 b) {
  uint4 r;
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}

constexpr int SPP_VECS = 4;   // 16-byte channel vectors per CTA

template <typename T>
__global__ void __launch_bounds__(256) spp_separable_kernel(const T* __restrict__ x, int x_ld, T* __restrict__ y, int y_ld, int h,
                                                            int w, int c) {
  constexpr int V = Vec16<T>::N;
  extern __shared__ uint4 spp_smem[];
  const int hw = h * w;
  uint4* cur = spp_smem;                 // [hw][SPP_VECS]
  uint4* tmp = spp_smem + hw * SPP_VECS;
  const int img = blockIdx.y;
  const int c0 = blockIdx.x * SPP_VECS * V;
  const T* xi = x + (long long)img * hw * x_ld + c0;
  T* yo = y + (long long)img * hw * y_ld + c0;
  const int items = hw * SPP_VECS;
  for (int i = threadIdx.x; i < items; i += blockDim.x) {
    const int pix = i / SPP_VECS, v = i % SPP_VECS;
    const uint4 val = __ldg(reinterpret_cast<const uint4*>(xi + (long long)pix * x_ld + v * V));
    cur[i] = val;
    *reinterpret_cast<uint4*>(yo + (long long)pix * y_ld + v * V) = val;
  }
  __syncthreads();
  for (int round = 1; round <= 3; ++round) {
    for (int i = threadIdx.x; i < items; i += blockDim.x) {      // row pass
      const int pix = i / SPP_VECS, v = i % SPP_VECS;
      const int px = pix % w, row0 = pix - px;
      uint4 m = cur[i];
      for (int d = -2; d <= 2; ++d) {
        const int xx = px + d;
        if (d != 0 && xx >= 0 && xx < w) m = vmax<T>(m, cur[(row0 + xx) * SPP_VECS + v]);
      }
      tmp[i] = m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < items; i += blockDim.x) {      // column pass
      const int pix = i / SPP_VECS, v = i % SPP_VECS;
      const int py = pix / w;
      uint4 m = tmp[i];
      for (int d = -2; d <= 2; ++d) {
        const int yy = py + d;
        if (d != 0 && yy >= 0 && yy < h) m = vmax<T>(m, tmp[(pix + d * w) * SPP_VECS + v]);
      }
      cur[i] = m;
      *reinterpret_cast<uint4*>(yo + (long long)pix * y_ld + round * c + v * V) = m;
    }
    __syncthreads();
  }
}

// SPP backward (training head): dx = dy[:, 0:c] + sum over the three pools of dy routed to each window's arg-max -- torch's
// max_pool2d backward: the FIRST maximum in row-major window order takes the gradient.  One CTA per (image, 32 channels): the input
// tile and an fp32 gradient tile live in shared memory.  The arg-max search is separable: per pool, a row pass leaves every pixel's
// first maximum of its horizontal window (value + column), a column pass scans those rows top to bottom with a strict '>' -- the
// first row holding the window maximum, and inside it the first column: exactly the row-major first maximum, for 2(2r+1) reads per
// item instead of (2r+1)^2 -- and adds the item's gradient with a shared-memory atomic (lane = channel: conflict-free); the tile is
// written once, no global atomics.
constexpr int SPPB_C = 32;    // channels per CTA
template <typename T>
__global__ void __launch_bounds__(1024) spp_backward_kernel(const T* __restrict__ x, int x_ld, const T* __restrict__ dy, int dy_ld,
                                                           T* __restrict__ dx, int dx_ld, int h, int w, int c) {
  extern __shared__ float sppb_smem[];
  const int hw = h * w;
  const int items = hw * SPPB_C;
  float* gs = sppb_smem;                                         // [hw][SPPB_C] gradient tile
  T* xs = reinterpret_cast<T*>(gs + items);                      // [hw][SPPB_C] input tile
  T* rv = xs + items;                                            // [hw][SPPB_C] row-pass maximum (one of the inputs: exact in T)
  unsigned char* ra = reinterpret_cast<unsigned char*>(rv + items);   // [hw][SPPB_C] row-pass arg-max column
  const int img = blockIdx.y, c0 = blockIdx.x * SPPB_C;
  const T* xi = x + (long long)img * hw * x_ld + c0;
  const T* dyi = dy + (long long)img * hw * dy_ld + c0;
  for (int i = threadIdx.x; i < items; i += blockDim.x) {
    const int pix = i / SPPB_C, ch = i % SPPB_C;
    xs[i] = xi[(long long)pix * x_ld + ch];
    gs[i] = to_f<T>(dyi[(long long)pix * dy_ld + ch]);            // identity branch of the concat
  }
  __syncthreads();
  for (int round = 1; round <= 3; ++round) {
    const int r = 2 * round;                                     // window radius: 2, 4, 6 (kernel 5, 9, 13; stride 1, same padding)
    for (int i = threadIdx.x; i < items; i += blockDim.x) {      // row pass
      const int pix = i / SPPB_C, ch = i % SPPB_C;
      const int px = pix % w, row0 = pix - px;
      const int x0 = max(px - r, 0), x1 = min(px + r, w - 1);
      float best = -CUDART_INF_F;
      int arg = x0;
      for (int xx = x0; xx <= x1; ++xx) {
        const float v = to_f<T>(xs[(row0 + xx) * SPPB_C + ch]);
        if (v > best) { best = v; arg = xx; }
      }
      rv[i] = xs[(row0 + arg) * SPPB_C + ch]; ra[i] = (unsigned char)arg;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < items; i += blockDim.x) {      // column pass + scatter
      const int pix = i / SPPB_C, ch = i % SPPB_C;
      const int py = pix / w, px = pix - py * w;
      const int y0 = max(py - r, 0), y1 = min(py + r, h - 1);
      float best = -CUDART_INF_F;
      int arg = y0 * w + ra[(y0 * w + px) * SPPB_C + ch];
      for (int yy = y0; yy <= y1; ++yy) {
        const int j = (yy * w + px) * SPPB_C + ch;
        const float v = to_f<T>(rv[j]);
        if (v > best) { best = v; arg = yy * w + ra[j]; }
      }
      atomicAdd(&gs[arg * SPPB_C + ch], to_f<T>(dyi[(long long)pix * dy_ld + round * c + ch]));
    }
    __syncthreads();
  }
  T* dxi = dx + (long long)img * hw * dx_ld + c0;
  for (int i = threadIdx.x; i < items; i += blockDim.x) dxi[(long long)(i / SPPB_C) * dx_ld + i % SPPB_C] = from_f<T>(gs[i]);
}

template <typename T>
__global__ void upsample2x_kernel(const T* __restrict__ x, int x_ld, T* __restrict__ y, int y_ld, int n, int h, int w, int c) {
  constexpr int V = Vec16<T>::N;
  const int cv = c / V;
  const int ho = 2 * h, wo = 2 * w;
  const long long total = (long long)n * ho * wo * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    long long p = i / cv;
    const int ox = (int)(p % wo); p /= wo;
    const int oy = (int)(p % ho);
    const int img = (int)(p / ho);
    uint4 raw = __ldg(reinterpret_cast<const uint4*>(x + (((long long)img * h + oy / 2) * w + ox / 2) * x_ld + v * V));
    *reinterpret_cast<uint4*>(y + (((long long)img * ho + oy) * wo + ox) * y_ld + v * V) = raw;
  }
}

template <typename T>
__global__ void copy_channels_kernel(const T* __restrict__ x, int x_ld, T* __restrict__ y, int y_ld, long long rows, int c) {
  constexpr int V = Vec16<T>::N;
  const int cv = c / V;
  const long long total = rows * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    const long long r = i / cv;
    *reinterpret_cast<uint4*>(y + r * y_ld + v * V) = __ldg(reinterpret_cast<const uint4*>(x + r * x_ld + v * V));
  }
}

template <typename T>
__global__ void coord_kernel(T* __restrict__ y, int ld, int h, int w) {
  const int total = h * w;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ix = i % w, iy = i / w;
    // arange / (w - 1) * 2.0 - 1 in fp32, custom_layers.py:267-268
    float cx = __fsub_rn(__fmul_rn(__fdiv_rn((float)ix, (float)(w - 1)), 2.f), 1.f);
    float cy = __fsub_rn(__fmul_rn(__fdiv_rn((float)iy, (float)(h - 1)), 2.f), 1.f);
    T* dst = y + (long long)i * ld;
    dst[0] = from_f<T>(cx);
    dst[1] = from_f<T>(cy);
    for (int k = 2; k < ld; ++k) dst[k] = from_f<T>(0.f);
  }
}

template <typename T>
__global__ void activation_kernel(T* __restrict__ x, long long count, int act) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    x[i] = from_f<T>(apply_act(to_f<T>(x[i]), act));
}

inline bool vec_ok(const void* p, int ld, int c, int dtype) {
  const int v = 16 / dtype_size(dtype);
  return p && (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % v == 0 && c % v == 0 && c > 0 && ld >= c;
}

}  // namespace
}  // namespace ppy

#define PPY_DISPATCH(dtype, ...)                                  \
  if ((dtype) == PPY_BF16) { using T = __nv_bfloat16; __VA_ARGS__ } \
  else if ((dtype) == PPY_F32) { using T = float; __VA_ARGS__ }     \
  else return PPY_ERR_INVALID;

extern "C" {
using namespace ppy;

int ppy_nchw_to_nhwc(const float* x, void* y, int n, int c, int h, int w, int y_ld, int y_dtype, ppy_stream_t s) {
  PPY_REQUIRE(x && y && n > 0 && c > 0 && h > 0 && w > 0 && y_ld >= c);
  const long long pixels = (long long)n * h * w;
  PPY_DISPATCH(y_dtype, nchw_to_nhwc_kernel<T><<<grid_for(pixels, 256), 256, 0, as_stream(s)>>>(x, (T*)y, n, c, h, w, y_ld);)
  return check_launch();
}

int ppy_nhwc_to_nchw(const void* x, int x_ld, int x_dtype, float* y, int n, int c, int h, int w, ppy_stream_t s) {
  PPY_REQUIRE(x && y && n > 0 && c > 0 && h > 0 && w > 0 && x_ld >= c);
  const long long total = (long long)n * c * h * w;
  PPY_DISPATCH(x_dtype, nhwc_to_nchw_kernel<T><<<grid_for(total, 256), 256, 0, as_stream(s)>>>((const T*)x, x_ld, y, n, c, h, w);)
  return check_launch();
}

int ppy_maxpool3x3s2(const void* x, int x_ld, void* y, int y_ld, int n, int h, int w, int c, int dtype, ppy_stream_t s) {
  PPY_REQUIRE(n > 0 && h > 0 && w > 0 && vec_ok(x, x_ld, c, dtype) && vec_ok(y, y_ld, c, dtype));
  const int ho = (h + 1) / 2, wo = (w + 1) / 2;
  const long long total = (long long)n * ho * wo * (c / (16 / dtype_size(dtype)));
  PPY_REQUIRE(total < 0x7FFFFFFFll);
  if (dtype == PPY_BF16) {
    const long long work = (long long)n * ho * ((wo + 1) / 2) * (c / 8);
    pool2_bf16_kernel<0><<<grid_for(work, 256), 256, 0, as_stream(s)>>>((const __nv_bfloat16*)x, x_ld, (__nv_bfloat16*)y, y_ld, n, h, w, c, ho, wo);
    return check_launch();
  }
  PPY_DISPATCH(dtype, pool_kernel<T, 0><<<grid_for(total, 256), 256, 0, as_stream(s)>>>((const T*)x, x_ld, (T*)y, y_ld, n, h, w, c, ho, wo);)
  return check_launch();
}

int ppy_avgpool2x2(const void* x, int x_ld, void* y, int y_ld, int n, int h, int w, int c, int dtype, ppy_stream_t s) {
  PPY_REQUIRE(n > 0 && h > 1 && w > 1 && vec_ok(x, x_ld, c, dtype) && vec_ok(y, y_ld, c, dtype));
  const int ho = h / 2, wo = w / 2;
  const long long total = (long long)n * ho * wo * (c / (16 / dtype_size(dtype)));
  PPY_REQUIRE(total < 0x7FFFFFFFll);
  if (dtype == PPY_BF16) {
    const long long work = (long long)n * ho * ((wo + 1) / 2) * (c / 8);
    pool2_bf16_kernel<1><<<grid_for(work, 256), 256, 0, as_stream(s)>>>((const __nv_bfloat16*)x, x_ld, (__nv_bfloat16*)y, y_ld, n, h, w, c, ho, wo);
    return check_launch();
  }
  PPY_DISPATCH(dtype, pool_kernel<T, 1><<<grid_for(total, 256), 256, 0, as_stream(s)>>>((const T*)x, x_ld, (T*)y, y_ld, n, h, w, c, ho, wo);)
  return check_launch();
}

int ppy_spp(const void* x, int x_ld, void* y, int y_ld, int n, int h, int w, int c, int dtype, ppy_stream_t s) {
  PPY_REQUIRE(n > 0 && h > 0 && w > 0 && vec_ok(x, x_ld, c, dtype) && vec_ok(y, y_ld, c, dtype) && y_ld >= 4 * c);
  const int vec = 16 / dtype_size(dtype);
  const size_t smem = (size_t)2 * h * w * SPP_VECS * 16;
  if (c % (SPP_VECS * vec) == 0 && smem <= 200 * 1024) {
    dim3 grid((unsigned)(c / (SPP_VECS * vec)), (unsigned)n);
    if (dtype == PPY_BF16) {
      if (smem > 48 * 1024) cudaFuncSetAttribute(spp_separable_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      spp_separable_kernel<__nv_bfloat16><<<grid, 256, smem, as_stream(s)>>>((const __nv_bfloat16*)x, x_ld, (__nv_bfloat16*)y, y_ld, h, w, c);
    } else if (dtype == PPY_F32) {
      if (smem > 48 * 1024) cudaFuncSetAttribute(spp_separable_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      spp_separable_kernel<float><<<grid, 256, smem, as_stream(s)>>>((const float*)x, x_ld, (float*)y, y_ld, h, w, c);
    } else return PPY_ERR_INVALID;
    return check_launch();
  }
  const long long total = (long long)n * h * w * (c / vec);
  PPY_DISPATCH(dtype, spp_kernel<T><<<grid_for(total, 128), 128, 0, as_stream(s)>>>((const T*)x, x_ld, (T*)y, y_ld, n, h, w, c);)
  return check_launch();
}

int ppy_spp_backward(const void* x, int x_ld, const void* dy, int dy_ld, void* dx, int dx_ld, int n, int h, int w, int c, int dtype,
                     ppy_stream_t s) {
  PPY_REQUIRE(x && dy && dx && n > 0 && h > 0 && w > 0 && c > 0 && c % SPPB_C == 0 && x_ld >= c && dx_ld >= c && dy_ld >= 4 * c);
  PPY_REQUIRE(dtype == PPY_BF16 || dtype == PPY_F32);
  const size_t smem = (size_t)h * w * SPPB_C * (4 + 2 * dtype_size(dtype) + 1);
  PPY_REQUIRE(smem <= 200 * 1024 && w <= 255);
  dim3 grid((unsigned)(c / SPPB_C), (unsigned)n);
  if (dtype == PPY_BF16) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(spp_backward_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    spp_backward_kernel<__nv_bfloat16><<<grid, 1024, smem, as_stream(s)>>>((const __nv_bfloat16*)x, x_ld, (const __nv_bfloat16*)dy, dy_ld,
                                                                          (__nv_bfloat16*)dx, dx_ld, h, w, c);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(spp_backward_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    spp_backward_kernel<float><<<grid, 1024, smem, as_stream(s)>>>((const float*)x, x_ld, (const float*)dy, dy_ld, (float*)dx, dx_ld, h, w, c);
  }
  return check_launch();
}

int ppy_upsample2x(const void* x, int x_ld, void* y, int y_ld, int n, int h, int w, int c, int dtype, ppy_stream_t s) {
  PPY_REQUIRE(n > 0 && h > 0 && w > 0 && vec_ok(x, x_ld, c, dtype) && vec_ok(y, y_ld, c, dtype));
  const long long total = (long long)n * h * w * 4 * (c / (16 / dtype_size(dtype)));
  PPY_DISPATCH(dtype, upsample2x_kernel<T><<<grid_for(total, 256), 256, 0, as_stream(s)>>>((const T*)x, x_ld, (T*)y, y_ld, n, h, w, c);)
  return check_launch();
}

int ppy_copy_channels(const void* x, int x_ld, void* y, int y_ld, long long rows, int c, int dtype, ppy_stream_t s) {
  PPY_REQUIRE(rows > 0 && vec_ok(x, x_ld, c, dtype) && vec_ok(y, y_ld, c, dtype));
  const long long total = rows * (c / (16 / dtype_size(dtype)));
  PPY_DISPATCH(dtype, copy_channels_kernel<T><<<grid_for(total, 256), 256, 0, as_stream(s)>>>((const T*)x, x_ld, (T*)y, y_ld, rows, c);)
  return check_launch();
}

int ppy_coord_channels(void* y, int y_ld, int h, int w, int dtype, ppy_stream_t s) {
  PPY_REQUIRE(y && y_ld >= 2 && h > 1 && w > 1);
  PPY_DISPATCH(dtype, coord_kernel<T><<<grid_for((long long)h * w, 128), 128, 0, as_stream(s)>>>((T*)y, y_ld, h, w);)
  return check_launch();
}

int ppy_activation(void* x, long long count, int act, int dtype, ppy_stream_t s) {
  PPY_REQUIRE(x && count > 0 && act >= PPY_ACT_NONE && act <= PPY_ACT_MISH);
  PPY_DISPATCH(dtype, activation_kernel<T><<<grid_for(count, 256), 256, 0, as_stream(s)>>>((T*)x, count, act);)
  return check_launch();
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// DCNv2 "gather" stage of the two-kernel deformable conv: one warp per (output pixel, tap) bilinearly samples
// the NHWC input at the learned offset, applies sigmoid(mask) and writes the C-channel run of
// A[m][tap*C + c] (bf16/fp32).  The matrix (M x 9C, e.g. 106 MB at bs=32 608^2) is consumed immediately by the
// TMA-fed 1x1 tcgen05 GEMM and stays L2-resident; the sampling math is identical to the fused kernel's.
// ------------------------------------------------------------------------------------------------
namespace ppy {
namespace {
template <typename T>
__global__ void __launch_bounds__(256) dcn_gather_kernel(const T* __restrict__ x, int x_ld, int n, int h, int w, int c,
                                                         const float* __restrict__ om, int om_ld, int k, int stride, int pad,
                                                         int ho, int wo, T* __restrict__ out) {
  constexpr int V = Vec16<T>::N;
  const int lane = threadIdx.x & 31;
  const int taps = k * k;
  const long long total = (long long)n * ho * wo * taps;
  const int cv = c / V;
  for (long long wi = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wi < total; wi += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int tap = (int)(wi % taps);
    const long long m = wi / taps;
    const int ox = (int)(m % wo), oy = (int)((m / wo) % ho), img = (int)(m / ((long long)wo * ho));
    const float* o = om + m * om_ld;
    const float dy = __ldg(o + 2 * tap), dx = __ldg(o + 2 * tap + 1), ml = __ldg(o + 2 * taps + tap);
    const float mask = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-ml)));
    const float py = (float)(oy * stride - pad + tap / k) + dy, px = (float)(ox * stride - pad + tap % k) + dx;
    const float fy = floorf(py), fx = floorf(px);
    const float ly = py - fy, lx = px - fx, hy = 1.f - ly, hx = 1.f - lx;
    const int y0 = (int)fminf(fmaxf(fy, -2.f), (float)h), x0 = (int)fminf(fmaxf(fx, -2.f), (float)w);
    const float wq[4] = {hy * hx * mask, hy * lx * mask, ly * hx * mask, ly * lx * mask};
    const T* src[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int yy = y0 + (q >> 1), xx = x0 + (q & 1);
      src[q] = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? x + (((long long)img * h + yy) * w + xx) * x_ld : nullptr;
    }
    T* dst = out + (m * taps + tap) * c;
    for (int v = lane; v < cv; v += 32) {
      float acc[V];
#pragma unroll
      for (int e = 0; e < V; ++e) acc[e] = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (src[q]) {
          float t[V];
          load_vec<T>(src[q] + v * V, t);
#pragma unroll
          for (int e = 0; e < V; ++e) acc[e] += wq[q] * t[e];
        }
      }
      store_vec<T>(dst + v * V, acc);
    }
  }
}

// k == 3 specialisation with the channel count fixed at compile time (C = ITERS * 32 lanes * 16-byte vectors): one warp per
// (output pixel, TPW taps).  The generic kernel above is bound by two chained L2 round trips per 1 KB of output (offset
// fetch, then one 4-corner load group per loop trip); here the warp's 3*TPW offset/mask values arrive in ONE coalesced
// load and are broadcast by shuffle, and all 4*ITERS corner loads of a tap are issued before the first blend, so the
// taps of a warp overlap each other's latency.  Out-of-image corners are loaded from a clamped in-image address with
// weight 0 (fma(0, t, acc) == acc), so results are bit-identical to the generic kernel.
template <typename T, int ITERS, int TPW>
__global__ void __launch_bounds__(256) dcn_gather3_kernel(const T* __restrict__ x, int x_ld, int n, int h, int w,
                                                          const float* __restrict__ om, int om_ld, int stride, int pad,
                                                          int ho, int wo, T* __restrict__ out) {
  constexpr int V = Vec16<T>::N;
  constexpr int C = ITERS * 32 * V;
  constexpr int GROUPS = 9 / TPW;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)n * ho * wo * GROUPS;
  for (long long wi = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wi < total; wi += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int grp = (int)(wi % GROUPS);
    const long long m = wi / GROUPS;
    const int ox = (int)(m % wo), oy = (int)((m / wo) % ho), img = (int)(m / ((long long)wo * ho));
    const float* o = om + m * om_ld;
    float val = 0.f;
    if (lane < 3 * TPW) {                        // lane = kind * TPW + j; kind 0: dy, 1: dx, 2: mask logit
      const int kind = lane / TPW, tap = grp * TPW + lane % TPW;
      val = __ldg(o + (kind == 2 ? 18 + tap : 2 * tap + kind));
    }
    const T* xi = x + (long long)img * h * w * x_ld;
    T* dst = out + (m * 9 + grp * TPW) * C;
#pragma unroll
    for (int j = 0; j < TPW; ++j) {
      const int tap = grp * TPW + j;
      const float dy = __shfl_sync(0xffffffffu, val, j), dx = __shfl_sync(0xffffffffu, val, TPW + j);
      const float ml = __shfl_sync(0xffffffffu, val, 2 * TPW + j);
      const float mask = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-ml)));
      const float py = (float)(oy * stride - pad + tap / 3) + dy, px = (float)(ox * stride - pad + tap % 3) + dx;
      const float fy = floorf(py), fx = floorf(px);
      const float ly = py - fy, lx = px - fx, hy = 1.f - ly, hx = 1.f - lx;
      const int y0 = (int)fminf(fmaxf(fy, -2.f), (float)h), x0 = (int)fminf(fmaxf(fx, -2.f), (float)w);
      float wq[4] = {hy * hx * mask, hy * lx * mask, ly * hx * mask, ly * lx * mask};
      const T* src[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int yy = y0 + (q >> 1), xx = x0 + (q & 1);
        const bool ok = yy >= 0 && yy < h && xx >= 0 && xx < w;
        if (!ok) wq[q] = 0.f;
        src[q] = xi + ((long long)min(max(yy, 0), h - 1) * w + min(max(xx, 0), w - 1)) * x_ld + lane * V;
      }
      uint4 raw[ITERS][4];
#pragma unroll
      for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int q = 0; q < 4; ++q) raw[it][q] = __ldg(reinterpret_cast<const uint4*>(src[q] + it * 32 * V));
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        float acc[V];
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const T* t = reinterpret_cast<const T*>(&raw[it][q]);
#pragma unroll
          for (int e = 0; e < V; ++e) acc[e] += wq[q] * to_f<T>(t[e]);
        }
        store_vec<T>(dst + j * C + (it * 32 + lane) * V, acc);
      }
    }
  }
}

template <typename T, int TPW>
bool launch_gather3(const T* x, int x_ld, int n, int h, int w, int c, const float* om, int om_ld, int stride, int pad, int ho,
                    int wo, T* out, cudaStream_t st) {
  constexpr int V = Vec16<T>::N;
  const long long warps = (long long)n * ho * wo * (9 / TPW);
  long long blocks = ceil_div(warps, 8);
  if (blocks > 148 * 64) blocks = 148 * 64;
#define PPY_G3(IT) dcn_gather3_kernel<T, IT, TPW><<<(unsigned)blocks, 256, 0, st>>>(x, x_ld, n, h, w, om, om_ld, stride, pad, ho, wo, out)
  if (c == 32 * V) PPY_G3(1);
  else if (c == 64 * V) PPY_G3(2);
  else if (c == 128 * V) PPY_G3(4);
  else return false;
#undef PPY_G3
  return true;
}
}  // namespace
}  // namespace ppy

extern "C" int ppy_dcn_gather(const void* x, int x_ld, int n, int h, int w, int c, const float* offset_mask, int om_ld, int k,
                              int stride, int pad, void* out, int dtype, ppy_stream_t s) {
  using namespace ppy;
  PPY_REQUIRE(x && offset_mask && out && n > 0 && h > 0 && w > 0 && k > 0 && stride > 0 && pad >= 0);
  PPY_REQUIRE(vec_ok(x, x_ld, c, dtype) && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && om_ld >= 3 * k * k);
  const int ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;
  PPY_REQUIRE(ho > 0 && wo > 0);
  if (k == 3) {   // compile-time-C specialisation, 3 taps per warp (75.8 -> 49.2 us at bs 32 x 19x19 x 512, bit-identical)
    bool done = false;
    PPY_DISPATCH(dtype, done = launch_gather3<T, 3>((const T*)x, x_ld, n, h, w, c, offset_mask, om_ld, stride, pad, ho, wo, (T*)out,
                                                   as_stream(s));)
    if (done) return check_launch();
  }
  const long long warps = (long long)n * ho * wo * k * k;
  long long blocks = ceil_div(warps, 8);
  if (blocks > 148 * 64) blocks = 148 * 64;
  PPY_DISPATCH(dtype, dcn_gather_kernel<T><<<(unsigned)blocks, 256, 0, as_stream(s)>>>((const T*)x, x_ld, n, h, w, c, offset_mask,
                                                                                     om_ld, k, stride, pad, ho, wo, (T*)out);)
  return check_launch();
}
