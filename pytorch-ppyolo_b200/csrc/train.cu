// Training-side kernels: train-mode BatchNorm for the (frozen) backbone -- the reference leaves its frozen BNs in train
// mode, so they normalise with BATCH statistics and keep updating the running ones (reference
// model/custom_layers.py:122, train.py:264) -- and the fused SGD-momentum update (reference train.py:437-442 with
// torch.optim.SGD semantics: weight decay folded into the gradient, dampening 0, no Nesterov).
//
//   bn_stats      per-channel sum / sum of squares over all pixels of an NHWC tensor (fp32 partials, fp64 atomics)
//   bn_finalize   mean / biased var -> folded scale, shift; running stats <- (1-m)*running + m*(mean, unbiased var)
//   scale_shift   y = act(x*scale[c] + shift[c] (+ residual)), 16-byte channel vectors
//   sgd_momentum  g' = g*grad_scale + wd*p ; buf = first ? g' : mu*buf + g' ; p -= lr*buf
//   bn_train_fused   statistics + finalize + normalise (+ residual, + activation) of one layer as ONE cooperative launch
//   bn_act_backward  activation' * dy, both BatchNorm reductions and dx of a head layer as ONE cooperative launch
//   kmajor           K-major (pixel-contiguous) operands of the weight-gradient GEMM: transposed im2col
//   sgd_ema_multi    optimizer step + EMA of all trainable tensors in one launch
#include <stdlib.h>
#include <cooperative_groups.h>
#include "common.cuh"

namespace ppy {
namespace {

template <typename T> struct VecT { static constexpr int N = 16 / sizeof(T); };

template <typename T>
__device__ __forceinline__ void ldv(const T* p, float (&v)[VecT<T>::N]) {
  uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
  const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
  for (int i = 0; i < VecT<T>::N; ++i) v[i] = to_f<T>(e[i]);
}

// Per-channel sum / sum of squares + the finalize step in ONE launch.  Block = LC channel-vector lanes x (256 / LC) row lanes with
// LC = min(32, channel vectors) -- a 64-channel bf16 tensor (8 vectors) keeps all 256 threads busy on 32 rows instead of 64 threads on
// 8 --, four independent 16-byte loads in flight per thread, fp32 partials per thread, fp64 atomics per block; the last block to
// finish (device counter behind the sums) turns the totals into folded scale / shift and updates the running statistics.
struct BnFinal {
  const float* gamma; const float* beta; float eps, momentum; float* running_mean; float* running_var; float* scale; float* shift;
  float* save_mean; float* save_invstd;      // optional: batch mean / 1/sqrt(var + eps) for a BatchNorm backward
};

__device__ __forceinline__ void bn_finalize_channel(const double* __restrict__ sums, long long rows, int c, int ch, const BnFinal& f) {
  const double mean = __ldcg(sums + ch) / (double)rows;
  double var = __ldcg(sums + c + ch) / (double)rows - mean * mean;
  if (var < 0.0) var = 0.0;
  const float inv = rsqrtf((float)var + f.eps) * (f.gamma ? f.gamma[ch] : 1.f);
  f.scale[ch] = inv;
  f.shift[ch] = (f.beta ? f.beta[ch] : 0.f) - (float)mean * inv;
  if (f.running_mean) f.running_mean[ch] = (1.f - f.momentum) * f.running_mean[ch] + f.momentum * (float)mean;
  if (f.running_var) {
    const double unbiased = rows > 1 ? var * (double)rows / (double)(rows - 1) : var;
    f.running_var[ch] = (1.f - f.momentum) * f.running_var[ch] + f.momentum * (float)unbiased;
  }
}

template <typename T, int LC>
__global__ void __launch_bounds__(256) bn_stats_kernel(const T* __restrict__ x, int ld, long long rows, int c,
                                                       double* __restrict__ sums, BnFinal fin) {
  constexpr int V = VecT<T>::N;
  constexpr int RL = 256 / LC;
  __shared__ float red[2][256][V];
  __shared__ bool last;
  const int vl = threadIdx.x % LC, rl = threadIdx.x / LC;
  const int vec = blockIdx.x * LC + vl;
  float s[V], q[V];
#pragma unroll
  for (int k = 0; k < V; ++k) { s[k] = 0.f; q[k] = 0.f; }
  if (vec * V < c) {
    const long long step = (long long)gridDim.y * RL;
    const T* col = x + vec * V;
    for (long long r0 = (long long)blockIdx.y * RL + rl; r0 < rows; r0 += 4 * step) {
      float v[4][V];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long r = r0 + u * step;
        if (r < rows) ldv<T>(col + r * ld, v[u]);
        else {
#pragma unroll
          for (int k = 0; k < V; ++k) v[u][k] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < V; ++k) { s[k] += v[u][k]; q[k] += v[u][k] * v[u][k]; }
    }
  }
#pragma unroll
  for (int k = 0; k < V; ++k) { red[0][threadIdx.x][k] = s[k]; red[1][threadIdx.x][k] = q[k]; }
  __syncthreads();
  if (rl == 0 && vec * V < c) {
#pragma unroll
    for (int k = 0; k < V; ++k) {
      double ds = 0.0, dq = 0.0;
      for (int j = 0; j < RL; ++j) { ds += red[0][j * LC + vl][k]; dq += red[1][j * LC + vl][k]; }
      atomicAdd(&sums[vec * V + k], ds);
      atomicAdd(&sums[c + vec * V + k], dq);
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* done = reinterpret_cast<unsigned int*>(sums + 2 * c);
    last = atomicAdd(done, 1u) == gridDim.x * gridDim.y - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  for (int ch = threadIdx.x; ch < c; ch += 256) bn_finalize_channel(sums, rows, c, ch, fin);
}

// Train-mode BatchNorm of one layer in ONE cooperative launch: batch statistics, finalize, running-stat update and the
// normalise + activation (+ residual) pass.  The backbone of a training step has 55 of these layers; as four graph nodes each
// (memset, statistics, finalize-in-last-block, apply) their launch ramps and tails cost more than their bytes (the launches move
// 1..280 MB).  Phase 1 = bn_stats_kernel's reduction (fp32 partials, fp64 atomics per CTA); grid-wide barrier; phase 2: every
// CTA derives scale / shift for all channels into shared memory and streams its share of the tensor (second read: mostly L2 hits,
// the conv has just written it).  The workspace is ZERO on entry and left zero on exit (the last CTA to have read the sums clears
// them): no memset node.  c <= kBnFusedMaxC.
constexpr int kBnFusedMaxC = 2048;
constexpr int kBnMaxReplicas = 16;      // workspace: kBnMaxReplicas * 2c + 1 doubles (PPY_BN_WORKSPACE_DOUBLES)

// CTA-level tail of a statistics phase: the 256 threads' fp32 partials (two sums per channel) are folded by ONE THREAD PER CHANNEL
// (LC*V threads, RL terms each, fp64) and added to replica `rep` of the global sums.  Replicas spread the CTAs of one channel column
// over several addresses: with 296 CTAs adding to the same 2c doubles, the same-address fp64 atomics (serialised in L2) were ~70 %
// of a small layer's launch (ncu: warps parked at the grid barrier waiting for them).
template <int LC, int V>
__device__ __forceinline__ void cta_fold_atomic(float (&red)[2][256][V], double* __restrict__ sums, int c, int cbase, int rep) {
  constexpr int RL = 256 / LC;
  const int t = threadIdx.x;
  if (t < LC * V && cbase + t < c) {
    const int vl = t / V, k = t % V;
    double ds = 0.0, dq = 0.0;
#pragma unroll 4
    for (int j = 0; j < RL; ++j) { ds += red[0][j * LC + vl][k]; dq += red[1][j * LC + vl][k]; }
    double* dst = sums + (size_t)rep * 2 * c;
    atomicAdd(dst + cbase + t, ds);
    atomicAdd(dst + c + cbase + t, dq);
  }
}

template <typename T, int LC>
__global__ void __launch_bounds__(256) bn_train_fused_kernel(const T* __restrict__ x, int x_ld, T* __restrict__ y, int y_ld, long long rows,
                                                             int c, double* __restrict__ sums, BnFinal fin, const T* __restrict__ res,
                                                             int res_ld, int act, int gx, int reps) {
  constexpr int V = VecT<T>::N;
  constexpr int RL = 256 / LC;
  __shared__ float red[2][256][V];
  __shared__ float s_scale[LC * V], s_shift[LC * V];
  const int bx = blockIdx.x % gx, by = blockIdx.x / gx, gy = gridDim.x / gx;      // 1-D grid (cooperative), viewed as gx x gy
  const int vl = threadIdx.x % LC, rl = threadIdx.x / LC;
  const int vec = bx * LC + vl;
  const int cbase = bx * LC * V;
  float s[V], q[V];
#pragma unroll
  for (int k = 0; k < V; ++k) { s[k] = 0.f; q[k] = 0.f; }
  if (vec * V < c) {
    const long long step = (long long)gy * RL;
    const T* col = x + vec * V;
    for (long long r0 = (long long)by * RL + rl; r0 < rows; r0 += 4 * step) {
      float v[4][V];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long r = r0 + u * step;
        if (r < rows) ldv<T>(col + r * x_ld, v[u]);
        else {
#pragma unroll
          for (int k = 0; k < V; ++k) v[u][k] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < V; ++k) { s[k] += v[u][k]; q[k] += v[u][k] * v[u][k]; }
    }
  }
#pragma unroll
  for (int k = 0; k < V; ++k) { red[0][threadIdx.x][k] = s[k]; red[1][threadIdx.x][k] = q[k]; }
  __syncthreads();
  cta_fold_atomic<LC, V>(red, sums, c, cbase, by % reps);
  __threadfence();
  cooperative_groups::this_grid().sync();
  // ---- phase 2: scale / shift of this CTA's channels (the by == 0 CTAs also publish them and move the running statistics)
  if (threadIdx.x < LC * V && cbase + threadIdx.x < c) {
    const int ch = cbase + threadIdx.x;
    double sm = 0.0, sq = 0.0;
    for (int r = 0; r < reps; ++r) { sm += __ldcg(sums + (size_t)r * 2 * c + ch); sq += __ldcg(sums + (size_t)r * 2 * c + c + ch); }
    const double mean = sm / (double)rows;
    double var = sq / (double)rows - mean * mean;
    if (var < 0.0) var = 0.0;
    const float inv = rsqrtf((float)var + fin.eps) * (fin.gamma ? fin.gamma[ch] : 1.f);
    const float sh = (fin.beta ? fin.beta[ch] : 0.f) - (float)mean * inv;
    s_scale[threadIdx.x] = inv; s_shift[threadIdx.x] = sh;
    if (by == 0) {
      fin.scale[ch] = inv; fin.shift[ch] = sh;
      if (fin.save_mean) fin.save_mean[ch] = (float)mean;
      if (fin.save_invstd) fin.save_invstd[ch] = rsqrtf((float)var + fin.eps);
      if (fin.running_mean) fin.running_mean[ch] = (1.f - fin.momentum) * fin.running_mean[ch] + fin.momentum * (float)mean;
      if (fin.running_var) {
        const double unbiased = rows > 1 ? var * (double)rows / (double)(rows - 1) : var;
        fin.running_var[ch] = (1.f - fin.momentum) * fin.running_var[ch] + fin.momentum * (float)unbiased;
      }
    }
  }
  __syncthreads();
  // the last CTA to have read the sums clears them (and the counter) for the next launch
  __shared__ bool last;
  if (threadIdx.x == 0) {
    unsigned int* done = reinterpret_cast<unsigned int*>(sums + (size_t)kBnMaxReplicas * 2 * c);
    last = atomicAdd(done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    for (int i = threadIdx.x; i < 2 * c * reps; i += 256) sums[i] = 0.0;
    if (threadIdx.x == 0) sums[(size_t)kBnMaxReplicas * 2 * c] = 0.0;
  }
  {
    // Same thread -> (channel vector, row lane) map as phase 1: the thread's V scales / shifts live in registers (no per-element
    // index division), four rows (+ residual rows) in flight per thread, and the sweep runs BACKWARDS over phase 1's row order
    // so the rows read last -- the ones still in L2 -- are re-read first.
    if (vec * V >= c) return;
    float sc[V], sf[V];
#pragma unroll
    for (int k = 0; k < V; ++k) { sc[k] = s_scale[vl * V + k]; sf[k] = s_shift[vl * V + k]; }
    const long long step = (long long)gy * RL, first = (long long)by * RL + rl;
    if (first >= rows) return;
    const T* xcol = x + vec * V;
    const T* rcol = res ? res + vec * V : nullptr;
    T* ycol = y + vec * V;
    for (long long r0 = first + (rows - 1 - first) / (4 * step) * (4 * step); r0 >= first; r0 -= 4 * step) {
      uint4 xa[4], ra[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long r = r0 + u * step;
        if (r < rows) {
          xa[u] = __ldg(reinterpret_cast<const uint4*>(xcol + r * x_ld));
          if (rcol) ra[u] = __ldg(reinterpret_cast<const uint4*>(rcol + r * res_ld));
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long r = r0 + u * step;
        if (r >= rows) continue;
        const T* e = reinterpret_cast<const T*>(&xa[u]);
        const T* f = reinterpret_cast<const T*>(&ra[u]);
        uint4 raw;
        T* o = reinterpret_cast<T*>(&raw);
#pragma unroll
        for (int k = 0; k < V; ++k) {
          float a = to_f<T>(e[k]) * sc[k] + sf[k];
          if (rcol) a += to_f<T>(f[k]);
          o[k] = from_f<T>(apply_act(a, act));
        }
        *reinterpret_cast<uint4*>(ycol + r * y_ld) = raw;
      }
    }
  }
}

// Backward of train-mode BatchNorm + relu / leaky(0.1) of a head layer in ONE cooperative launch (ATen: activation backward +
// batch_norm_backward_reduce + batch_norm_backward_elemt, three launches and 6 reads + 2 writes of small L2-resident tensors).
//   g = dy * act'(y);  dbeta = sum g;  dgamma = sum g * xhat  (xhat = (x - mean) * invstd)
//   dx = gamma * invstd * (g - dbeta / N - xhat * dgamma / N)
// Phase 1: per-thread fp32 partials of the two sums over the thread's V channels (bn_train_fused's thread map), fp64 atomics per
// CTA (replicated sums, cta_fold_atomic); grid barrier; phase 2 re-reads dy / y / x (L2 hits) backwards and writes dx.  Workspace:
// PPY_BN_WORKSPACE_DOUBLES(c), zero on entry and left zero.
template <typename T, int LC>
__global__ void __launch_bounds__(256) bn_act_backward_kernel(const T* __restrict__ dy, int dy_ld, const T* __restrict__ x, int x_ld,
                                                              const T* __restrict__ y, int y_ld, T* __restrict__ dx, int dx_ld, long long rows,
                                                              int c, const float* __restrict__ gamma, const float* __restrict__ mean,
                                                              const float* __restrict__ invstd, int act, float* __restrict__ dgamma,
                                                              float* __restrict__ dbeta, double* __restrict__ sums, int gx, int reps) {
  constexpr int V = VecT<T>::N;
  constexpr int RL = 256 / LC;
  __shared__ float red[2][256][V];
  const int bx = blockIdx.x % gx, by = blockIdx.x / gx, gy = gridDim.x / gx;
  const int vl = threadIdx.x % LC, rl = threadIdx.x / LC;
  const int vec = bx * LC + vl;
  const int cbase = bx * LC * V;
  const bool live = vec * V < c;
  const float slope = act == PPY_ACT_RELU ? 0.f : act == PPY_ACT_LEAKY ? 0.1f : 1.f;
  const bool masked = act != PPY_ACT_NONE && y != nullptr;
  float mu[V], is[V];
#pragma unroll
  for (int k = 0; k < V; ++k) { mu[k] = live ? mean[vec * V + k] : 0.f; is[k] = live ? invstd[vec * V + k] : 0.f; }
  float s[V], q[V];
#pragma unroll
  for (int k = 0; k < V; ++k) { s[k] = 0.f; q[k] = 0.f; }
  const long long step = (long long)gy * RL, first = (long long)by * RL + rl;
  const T* dcol = dy + vec * V;
  const T* xcol = x + vec * V;
  const T* ycol = masked ? y + vec * V : nullptr;
  if (live) {
    for (long long r0 = first; r0 < rows; r0 += 2 * step) {
      uint4 da[2], xa[2], ya[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const long long r = r0 + u * step;
        if (r < rows) {
          da[u] = __ldg(reinterpret_cast<const uint4*>(dcol + r * dy_ld));
          xa[u] = __ldg(reinterpret_cast<const uint4*>(xcol + r * x_ld));
          if (masked) ya[u] = __ldg(reinterpret_cast<const uint4*>(ycol + r * y_ld));
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (r0 + u * step >= rows) continue;
        const T* de = reinterpret_cast<const T*>(&da[u]);
        const T* xe = reinterpret_cast<const T*>(&xa[u]);
        const T* ye = reinterpret_cast<const T*>(&ya[u]);
#pragma unroll
        for (int k = 0; k < V; ++k) {
          float g = to_f<T>(de[k]);
          if (masked && !(to_f<T>(ye[k]) > 0.f)) g *= slope;
          s[k] += g;
          q[k] += g * ((to_f<T>(xe[k]) - mu[k]) * is[k]);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < V; ++k) { red[0][threadIdx.x][k] = s[k]; red[1][threadIdx.x][k] = q[k]; }
  __syncthreads();
  cta_fold_atomic<LC, V>(red, sums, c, cbase, by % reps);
  __threadfence();
  cooperative_groups::this_grid().sync();
  __shared__ float s_k1[LC * V], s_k2[LC * V];
  if (threadIdx.x < LC * V && cbase + threadIdx.x < c) {
    const int ch = cbase + threadIdx.x;
    double sg = 0.0, sq = 0.0;
    for (int r = 0; r < reps; ++r) { sg += __ldcg(sums + (size_t)r * 2 * c + ch); sq += __ldcg(sums + (size_t)r * 2 * c + c + ch); }
    s_k1[threadIdx.x] = (float)(sg / (double)rows); s_k2[threadIdx.x] = (float)(sq / (double)rows);
    if (by == 0) {
      if (dbeta) dbeta[ch] = (float)sg;
      if (dgamma) dgamma[ch] = (float)sq;
    }
  }
  __syncthreads();
  float k1[V], k2[V], gi[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    k1[k] = live ? s_k1[vl * V + k] : 0.f; k2[k] = live ? s_k2[vl * V + k] : 0.f;
    gi[k] = live ? (gamma ? gamma[vec * V + k] : 1.f) * is[k] : 0.f;
  }
  __shared__ bool last;
  if (threadIdx.x == 0) {
    unsigned int* done = reinterpret_cast<unsigned int*>(sums + (size_t)kBnMaxReplicas * 2 * c);
    last = atomicAdd(done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    for (int i = threadIdx.x; i < 2 * c * reps; i += 256) sums[i] = 0.0;
    if (threadIdx.x == 0) sums[(size_t)kBnMaxReplicas * 2 * c] = 0.0;
  }
  if (!live || first >= rows) return;
  T* ocol = dx + vec * V;
  for (long long r0 = first + (rows - 1 - first) / (2 * step) * (2 * step); r0 >= first; r0 -= 2 * step) {
    uint4 da[2], xa[2], ya[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long r = r0 + u * step;
      if (r < rows) {
        da[u] = __ldg(reinterpret_cast<const uint4*>(dcol + r * dy_ld));
        xa[u] = __ldg(reinterpret_cast<const uint4*>(xcol + r * x_ld));
        if (masked) ya[u] = __ldg(reinterpret_cast<const uint4*>(ycol + r * y_ld));
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long r = r0 + u * step;
      if (r >= rows) continue;
      const T* de = reinterpret_cast<const T*>(&da[u]);
      const T* xe = reinterpret_cast<const T*>(&xa[u]);
      const T* ye = reinterpret_cast<const T*>(&ya[u]);
      uint4 raw;
      T* o = reinterpret_cast<T*>(&raw);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        float g = to_f<T>(de[k]);
        if (masked && !(to_f<T>(ye[k]) > 0.f)) g *= slope;
        const float xh = (to_f<T>(xe[k]) - mu[k]) * is[k];
        o[k] = from_f<T>((g - k1[k] - xh * k2[k]) * gi[k]);
      }
      *reinterpret_cast<uint4*>(ocol + r * dx_ld) = raw;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) scale_shift_act_kernel(const T* __restrict__ x, int x_ld, T* __restrict__ y, int y_ld,
                                                              long long rows, int c, const float* __restrict__ scale,
                                                              const float* __restrict__ shift, const T* __restrict__ res,
                                                              int res_ld, int act) {
  // (four vectors in flight per thread were measured: 1.25 -> 1.64 ms over the 55 backbone layers of a bs-8 step -- the launches
  // are small, 1..280 MB, and bound by their ramp, not by loads in flight)
  constexpr int V = VecT<T>::N;
  const int cv = c / V;
  const long long total = rows * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    const long long r = i / cv;
    float a[V];
    ldv<T>(x + r * x_ld + v * V, a);
    if (res) {
      float b[V];
      ldv<T>(res + r * res_ld + v * V, b);
#pragma unroll
      for (int k = 0; k < V; ++k) a[k] = a[k] * __ldg(scale + v * V + k) + __ldg(shift + v * V + k) + b[k];
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) a[k] = a[k] * __ldg(scale + v * V + k) + __ldg(shift + v * V + k);
    }
    uint4 raw;
    T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
    for (int k = 0; k < V; ++k) e[k] = from_f<T>(apply_act(a[k], act));
    *reinterpret_cast<uint4*>(y + r * y_ld + v * V) = raw;
  }
}

__global__ void sgd_momentum_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, long long n,
                                    float lr, float momentum, float wd, float grad_scale, int first_step) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float w = p[i];
    const float d = g[i] * grad_scale + wd * w;
    const float b = first_step ? d : momentum * buf[i] + d;
    buf[i] = b;
    p[i] = w - lr * b;
  }
}

inline unsigned blocks_for(long long work, int threads, long long cap = 148ll * 16) {
  long long b = ceil_div(work, threads);
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

// K-major operand builder of the weight-gradient GEMM (conv_autograd.py): out[(c*taps + tap)][m] = x[pixel m shifted by tap][c]
// (zero outside the image and for m in [M, m_pad)), i.e. the transposed im2col matrix with rows in the weight's OIHW order, or
// the plain transpose for k = 1.  One 64-pixel x 64-channel tile per CTA and tap, transposed through shared memory: reads are
// 128-byte channel runs of NHWC pixels, writes 128-byte pixel runs of one output row.
__global__ void __launch_bounds__(256) kmajor_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int n, int h, int w, int c, int k, int pad,
                                                     int stride, int ho, int wo, __nv_bfloat16* __restrict__ out, long long m_pad) {
  // tile[pixel row][16-byte chunk of 8 channels], chunk index XOR-swizzled with (row / 8): the 16-byte row stores and the 32-bit
  // column reads of the transpose are both bank-conflict free (the first version moved every element with 2-byte shared-memory
  // accesses: 32 per thread, instruction-bound at ~1.6 TB/s of writes)
  __shared__ uint4 tile[64][8];
  const long long m0 = (long long)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64, tap = blockIdx.z, ky = tap / k, kx = tap % k;
  const long long M = (long long)n * ho * wo;          // columns = OUTPUT pixels (stride 1: the input's own grid)
  // load: thread -> (pixel row r, 8 channels = one 16-byte vector), two passes
  for (int e = threadIdx.x; e < 64 * 8; e += 256) {
    const int r = e >> 3, v = e & 7;
    const long long m = m0 + r;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (m < M && c0 + v * 8 < c) {
      // (32-bit index arithmetic: M < 2^31 is checked by the caller; three 64-bit divisions per vector were most of the kernel)
      const unsigned mu = (unsigned)m, rowi = mu / (unsigned)wo;
      const int px = (int)(mu - rowi * (unsigned)wo), py = (int)(rowi % (unsigned)ho);
      const long long img = rowi / (unsigned)ho;
      const int iy = py * stride + ky - pad, ix = px * stride + kx - pad;
      if (iy >= 0 && iy < h && ix >= 0 && ix < w)
        val = __ldg(reinterpret_cast<const uint4*>(x + ((img * h + iy) * w + ix) * x_ld + c0 + v * 8));
    }
    tile[r][v ^ ((r >> 3) & 7)] = val;
  }
  __syncthreads();
  // store: thread -> (channel PAIR cp, eight consecutive pixels 8v .. 8v+7): eight 32-bit reads (two channels of one pixel each),
  // byte-permuted into the two channels' 16-byte pixel runs
  const int taps = k * k;
  const int cp = threadIdx.x >> 3, v = threadIdx.x & 7;
  const int chunk = cp >> 2, wi = cp & 3;
  uint32_t q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) q[j] = reinterpret_cast<const uint32_t*>(&tile[v * 8 + j][chunk ^ v])[wi];
  if (m0 + v * 8 >= m_pad) return;
  const int cc = c0 + 2 * cp;
  if (cc < c) {
    const uint4 lo = make_uint4(__byte_perm(q[0], q[1], 0x5410), __byte_perm(q[2], q[3], 0x5410), __byte_perm(q[4], q[5], 0x5410),
                                __byte_perm(q[6], q[7], 0x5410));
    *reinterpret_cast<uint4*>(out + ((long long)cc * taps + tap) * m_pad + m0 + v * 8) = lo;
  }
  if (cc + 1 < c) {
    const uint4 hi = make_uint4(__byte_perm(q[0], q[1], 0x7632), __byte_perm(q[2], q[3], 0x7632), __byte_perm(q[4], q[5], 0x7632),
                                __byte_perm(q[6], q[7], 0x7632));
    *reinterpret_cast<uint4*>(out + ((long long)(cc + 1) * taps + tap) * m_pad + m0 + v * 8) = hi;
  }
}

// blockIdx.y = tensor, blockIdx.x strides its elements; offsets[t] .. offsets[t+1] is tensor t's range in the flat shadow
__global__ void __launch_bounds__(256) ema_update_kernel(float* __restrict__ shadow, const float* const* __restrict__ params,
                                                         const long long* __restrict__ offsets, float decay, float one_minus_decay) {
  const int t = blockIdx.y;
  const long long begin = offsets[t], count = offsets[t + 1] - begin;
  const float* __restrict__ p = params[t];
  float* __restrict__ sh = shadow + begin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    sh[i] = __fadd_rn(__fmul_rn(decay, sh[i]), __fmul_rn(one_minus_decay, p[i]));
}

// Whole optimizer step in ONE launch (train.py:437-442 + model/EMA.py:31-45): for every trainable tensor t (blockIdx.y) the
// torch.optim.SGD momentum update from the all-reduced flat gradient bucket, then -- when shadow != nullptr -- the EMA of the
// updated value, in the operation order of the two separate kernels above (results are bit-identical to running them one after
// the other).  offsets: tensor ranges in the flat gradient / momentum buffers; shadow_offsets: start of tensor t in the EMA's
// flat shadow (its own tensor order); lr_mult / wd: per tensor (the reference's per-layer parameter groups).
__global__ void __launch_bounds__(256) sgd_ema_multi_kernel(float* const* __restrict__ params, const float* __restrict__ grad,
                                                            float* __restrict__ mom, float* __restrict__ shadow,
                                                            const long long* __restrict__ offsets, const long long* __restrict__ shadow_offsets,
                                                            const float* __restrict__ lr_mult, const float* __restrict__ wd, float lr,
                                                            float momentum, float grad_scale, int first_step, float decay,
                                                            float one_minus_decay) {
  const int t = blockIdx.y;
  const long long begin = offsets[t], count = offsets[t + 1] - begin;
  float* __restrict__ p = params[t];
  const float* __restrict__ g = grad + begin;
  float* __restrict__ buf = mom + begin;
  float* __restrict__ sh = shadow ? shadow + shadow_offsets[t] : nullptr;
  const float lr_t = lr * lr_mult[t], wd_t = wd[t];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < count; i0 += 4 * stride) {
    float w[4], gr[4], bu[4], so[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {                      // four independent elements in flight per thread and array
      const long long i = i0 + u * stride;
      if (i < count) { w[u] = p[i]; gr[u] = g[i]; bu[u] = first_step ? 0.f : buf[i]; so[u] = sh ? sh[i] : 0.f; }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      if (i >= count) continue;
      const float d = gr[u] * grad_scale + wd_t * w[u];
      const float b = first_step ? d : momentum * bu[u] + d;
      buf[i] = b;
      const float w2 = w[u] - lr_t * b;
      p[i] = w2;
      if (sh) sh[i] = __fadd_rn(__fmul_rn(decay, so[u]), __fmul_rn(one_minus_decay, w2));
    }
  }
}

// replicas of the per-channel sums: ~16 CTAs per address, at most kBnMaxReplicas (PPY_BN_REPLICAS overrides: experiments)
inline int bn_replicas(long long gy) {
  static const char* knob = getenv("PPY_BN_REPLICAS");
  int r = knob ? atoi(knob) : (int)(gy / 16);
  if (r > kBnMaxReplicas) r = kBnMaxReplicas;
  return r < 1 ? 1 : r;
}

template <typename T, int LC>
static int launch_bn_fused(const void* x, int x_ld, void* y, int y_ld, long long rows, int c, double* ws, const BnFinal& fin, const void* res,
                           int res_ld, int act, int gx, cudaStream_t st) {
  int dev = 0, sms = 0, per_sm = 0;
  int rc = check_cuda(cudaGetDevice(&dev));
  if (rc) return rc;
  if ((rc = check_cuda(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)))) return rc;
  if ((rc = check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bn_train_fused_kernel<T, LC>, 256, 0)))) return rc;
  if (per_sm < 1) return PPY_ERR_UNSUPPORTED;
  if (per_sm > 4) per_sm = 4;
  const int rl = 256 / LC;
  long long gy = ceil_div(rows, (long long)rl * 8);
  const long long cap = (long long)sms * per_sm / gx;          // every CTA resident (grid barrier)
  if (cap < 1) return PPY_ERR_UNSUPPORTED;
  if (gy > cap) gy = cap;
  if (gy < 1) gy = 1;
  const T* xp = (const T*)x; T* yp = (T*)y; const T* rp = (const T*)res;
  BnFinal f = fin;
  int reps = bn_replicas(gy);
  void* args[] = {(void*)&xp, (void*)&x_ld, (void*)&yp, (void*)&y_ld, (void*)&rows, (void*)&c, (void*)&ws, (void*)&f, (void*)&rp,
                  (void*)&res_ld, (void*)&act, (void*)&gx, (void*)&reps};
  rc = check_cuda(cudaLaunchCooperativeKernel((const void*)bn_train_fused_kernel<T, LC>, dim3((unsigned)(gx * gy)), dim3(256), args, 0, st));
  if (rc) return rc;
  count_launch();
  return PPY_OK;
}

template <typename T, int LC>
static int launch_bn_act_backward(const void* dy, int dy_ld, const void* x, int x_ld, const void* y, int y_ld, void* dx, int dx_ld, long long rows,
                                  int c, const float* gamma, const float* mean, const float* invstd, int act, float* dgamma, float* dbeta,
                                  double* ws, int gx, cudaStream_t st) {
  int dev = 0, sms = 0, per_sm = 0;
  int rc = check_cuda(cudaGetDevice(&dev));
  if (rc) return rc;
  if ((rc = check_cuda(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)))) return rc;
  if ((rc = check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bn_act_backward_kernel<T, LC>, 256, 0)))) return rc;
  if (per_sm < 1) return PPY_ERR_UNSUPPORTED;
  if (per_sm > 4) per_sm = 4;
  const int rl = 256 / LC;
  long long gy = ceil_div(rows, (long long)rl * 4);
  const long long cap = (long long)sms * per_sm / gx;          // every CTA resident (grid barrier)
  if (cap < 1) return PPY_ERR_UNSUPPORTED;
  if (gy > cap) gy = cap;
  if (gy < 1) gy = 1;
  const T* dyp = (const T*)dy; const T* xp = (const T*)x; const T* yp = (const T*)y; T* dxp = (T*)dx;
  int reps = bn_replicas(gy);
  void* args[] = {(void*)&dyp, (void*)&dy_ld, (void*)&xp, (void*)&x_ld, (void*)&yp, (void*)&y_ld, (void*)&dxp, (void*)&dx_ld, (void*)&rows,
                  (void*)&c, (void*)&gamma, (void*)&mean, (void*)&invstd, (void*)&act, (void*)&dgamma, (void*)&dbeta, (void*)&ws, (void*)&gx,
                  (void*)&reps};
  rc = check_cuda(cudaLaunchCooperativeKernel((const void*)bn_act_backward_kernel<T, LC>, dim3((unsigned)(gx * gy)), dim3(256), args, 0, st));
  if (rc) return rc;
  count_launch();
  return PPY_OK;
}

}  // namespace
}  // namespace ppy

extern "C" {
using namespace ppy;

int ppy_sgd_ema_multi(float* const* params, const float* grad_flat, float* momentum_flat, float* shadow_flat, const long long* offsets,
                      const long long* shadow_offsets, const float* lr_mult, const float* weight_decay, int num_tensors, float lr,
                      float momentum, float grad_scale, int first_step, float ema_decay, float ema_one_minus_decay, ppy_stream_t s) {
  PPY_REQUIRE(params && grad_flat && momentum_flat && offsets && lr_mult && weight_decay && num_tensors > 0);
  PPY_REQUIRE(!shadow_flat || shadow_offsets);
  sgd_ema_multi_kernel<<<dim3(128, (unsigned)num_tensors), 256, 0, as_stream(s)>>>(params, grad_flat, momentum_flat, shadow_flat, offsets,
                                                                                  shadow_offsets, lr_mult, weight_decay, lr, momentum,
                                                                                  grad_scale, first_step, ema_decay, ema_one_minus_decay);
  return check_launch();
}

int ppy_bn_batch_stats(const void* x, int x_ld, long long rows, int c, int dtype, const float* gamma, const float* beta,
                       float eps, float momentum, float* running_mean, float* running_var, float* scale, float* shift,
                       double* workspace /* 2*c + 1 doubles */, ppy_stream_t s) {
  PPY_REQUIRE(x && scale && shift && workspace && rows > 0 && c > 0 && x_ld >= c);
  const int v = 16 / dtype_size(dtype);
  PPY_REQUIRE(c % v == 0 && x_ld % v == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0);
  PPY_REQUIRE(dtype == PPY_BF16 || dtype == PPY_F32);
  cudaStream_t st = as_stream(s);
  int rc = check_cuda(cudaMemsetAsync(workspace, 0, sizeof(double) * (2 * (size_t)c + 1), st));
  if (rc) return rc;
  const int cv = c / v;
  int lc = 32;
  while (lc > 1 && lc / 2 >= cv) lc >>= 1;           // smallest power of two >= min(cv, 32)
  const int rl = 256 / lc;
  const int gx = (int)ceil_div(cv, lc);
  long long gy = ceil_div(rows, (long long)rl * 8);    // >= 8 rows per row-lane
  // two CTAs per SM at most: every CTA ends with 2 fp64 atomics per channel on the same addresses, and with 8 CTAs per SM those
  // serialised atomics cost more than the loads (55 backbone layers of a bs-8 step: 1.58 -> 1.14 ms; 1 per SM: the same)
  const long long cap = ceil_div(148ll * 2, gx);
  if (gy > cap) gy = cap;
  if (gy < 1) gy = 1;
  dim3 grid((unsigned)gx, (unsigned)gy);
  const BnFinal fin = {gamma, beta, eps, momentum, running_mean, running_var, scale, shift, nullptr, nullptr};
#define PPY_BN(T, LC) bn_stats_kernel<T, LC><<<grid, 256, 0, st>>>((const T*)x, x_ld, rows, c, workspace, fin)
#define PPY_BN_LC(T) do { switch (lc) { case 32: PPY_BN(T, 32); break; case 16: PPY_BN(T, 16); break; case 8: PPY_BN(T, 8); break; \
                                       case 4: PPY_BN(T, 4); break; case 2: PPY_BN(T, 2); break; default: PPY_BN(T, 1); } } while (0)
  if (dtype == PPY_BF16) PPY_BN_LC(__nv_bfloat16);
  else PPY_BN_LC(float);
#undef PPY_BN_LC
#undef PPY_BN
  return check_launch();
}

int ppy_bn_train_fused(const void* x, int x_ld, void* y, int y_ld, long long rows, int c, int dtype, const float* gamma, const float* beta,
                       float eps, float momentum, float* running_mean, float* running_var, float* scale, float* shift,
                       const void* residual, int res_ld, int act, double* workspace /* PPY_BN_WORKSPACE_DOUBLES(c), ZERO on entry, zero on exit */,
                       float* save_mean, float* save_invstd, ppy_stream_t s) {
  PPY_REQUIRE(x && y && scale && shift && workspace && rows > 0 && c > 0 && c <= kBnFusedMaxC && x_ld >= c && y_ld >= c);
  PPY_REQUIRE(dtype == PPY_BF16 || dtype == PPY_F32);
  const int v = 16 / dtype_size(dtype);
  PPY_REQUIRE(c % v == 0 && x_ld % v == 0 && y_ld % v == 0 && (!residual || res_ld % v == 0));
  PPY_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(residual)) & 15) == 0);
  PPY_REQUIRE(act >= PPY_ACT_NONE && act <= PPY_ACT_MISH);
  const int cv = c / v;
  int lc = 32;
  while (lc > 1 && lc / 2 >= cv) lc >>= 1;
  const int gx = (int)ceil_div(cv, lc);
  const BnFinal fin = {gamma, beta, eps, momentum, running_mean, running_var, scale, shift, save_mean, save_invstd};
  cudaStream_t st = as_stream(s);
#define PPY_BNF(T, LC) return launch_bn_fused<T, LC>(x, x_ld, y, y_ld, rows, c, workspace, fin, residual, res_ld, act, gx, st)
#define PPY_BNF_LC(T) do { switch (lc) { case 32: PPY_BNF(T, 32); case 16: PPY_BNF(T, 16); case 8: PPY_BNF(T, 8); case 4: PPY_BNF(T, 4); \
                                        case 2: PPY_BNF(T, 2); default: PPY_BNF(T, 1); } } while (0)
  if (dtype == PPY_BF16) PPY_BNF_LC(__nv_bfloat16);
  PPY_BNF_LC(float);
#undef PPY_BNF_LC
#undef PPY_BNF
}

int ppy_bn_act_backward(const void* dy, int dy_ld, const void* x, int x_ld, const void* y, int y_ld, void* dx, int dx_ld, long long rows, int c,
                        int dtype, const float* gamma, const float* save_mean, const float* save_invstd, int act, float* dgamma, float* dbeta,
                        double* workspace /* PPY_BN_WORKSPACE_DOUBLES(c), ZERO on entry, zero on exit */, ppy_stream_t s) {
  PPY_REQUIRE(dy && x && dx && save_mean && save_invstd && workspace && rows > 0 && c > 0 && dy_ld >= c && x_ld >= c && dx_ld >= c);
  PPY_REQUIRE(dtype == PPY_BF16 || dtype == PPY_F32);
  PPY_REQUIRE(act == PPY_ACT_NONE || ((act == PPY_ACT_RELU || act == PPY_ACT_LEAKY) && y && y_ld >= c));
  const int v = 16 / dtype_size(dtype);
  PPY_REQUIRE(c % v == 0 && dy_ld % v == 0 && x_ld % v == 0 && dx_ld % v == 0 && (!y || y_ld % v == 0));
  PPY_REQUIRE(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0);
  const int cv = c / v;
  int lc = 32;
  while (lc > 1 && lc / 2 >= cv) lc >>= 1;
  const int gx = (int)ceil_div(cv, lc);
  cudaStream_t st = as_stream(s);
#define PPY_BNB(T, LC) return launch_bn_act_backward<T, LC>(dy, dy_ld, x, x_ld, y, y_ld, dx, dx_ld, rows, c, gamma, save_mean, save_invstd, act, \
                                                            dgamma, dbeta, workspace, gx, st)
#define PPY_BNB_LC(T) do { switch (lc) { case 32: PPY_BNB(T, 32); case 16: PPY_BNB(T, 16); case 8: PPY_BNB(T, 8); case 4: PPY_BNB(T, 4); \
                                        case 2: PPY_BNB(T, 2); default: PPY_BNB(T, 1); } } while (0)
  if (dtype == PPY_BF16) PPY_BNB_LC(__nv_bfloat16);
  PPY_BNB_LC(float);
#undef PPY_BNB_LC
#undef PPY_BNB
}

int ppy_scale_shift_act(const void* x, int x_ld, void* y, int y_ld, long long rows, int c, int dtype, const float* scale,
                        const float* shift, const void* residual, int res_ld, int act, ppy_stream_t s) {
  PPY_REQUIRE(x && y && scale && shift && rows > 0 && c > 0 && x_ld >= c && y_ld >= c);
  const int v = 16 / dtype_size(dtype);
  PPY_REQUIRE(c % v == 0 && x_ld % v == 0 && y_ld % v == 0 && (!residual || res_ld % v == 0));
  PPY_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(residual)) & 15) == 0);
  PPY_REQUIRE(act >= PPY_ACT_NONE && act <= PPY_ACT_MISH);
  const long long total = rows * (c / v);
  if (dtype == PPY_BF16)
    scale_shift_act_kernel<__nv_bfloat16><<<blocks_for(total, 256, 148ll * 32), 256, 0, as_stream(s)>>>(
        (const __nv_bfloat16*)x, x_ld, (__nv_bfloat16*)y, y_ld, rows, c, scale, shift, (const __nv_bfloat16*)residual, res_ld, act);
  else if (dtype == PPY_F32)
    scale_shift_act_kernel<float><<<blocks_for(total, 256, 148ll * 32), 256, 0, as_stream(s)>>>(
        (const float*)x, x_ld, (float*)y, y_ld, rows, c, scale, shift, (const float*)residual, res_ld, act);
  else return PPY_ERR_INVALID;
  return check_launch();
}

int ppy_sgd_momentum(float* param, const float* grad, float* momentum_buf, long long n, float lr, float momentum,
                     float weight_decay, float grad_scale, int first_step, ppy_stream_t s) {
  PPY_REQUIRE(param && grad && momentum_buf && n > 0);
  sgd_momentum_kernel<<<blocks_for(n, 256), 256, 0, as_stream(s)>>>(param, grad, momentum_buf, n, lr, momentum, weight_decay,
                                                                    grad_scale, first_step);
  return check_launch();
}

int ppy_im2col_kmajor_strided(const void* x, int x_ld, int n, int h, int w, int c, int k, int stride, int pad, void* out, long long m_pad,
                              ppy_stream_t s) {
  PPY_REQUIRE(x && out && n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0 && x_ld >= c && (x_ld * 2) % 16 == 0 && k >= 1 && k <= 7 && pad >= 0);
  PPY_REQUIRE(stride >= 1 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0);
  const int ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;
  PPY_REQUIRE(ho > 0 && wo > 0);
  const long long M = (long long)n * ho * wo;
  PPY_REQUIRE(m_pad >= M && m_pad % 64 == 0 && m_pad < 0x7FFFFFFFll);      // (the kernel indexes pixels with 32 bits)
  dim3 grid((unsigned)(m_pad / 64), (unsigned)ceil_div(c, 64), (unsigned)(k * k));
  kmajor_kernel<<<grid, 256, 0, as_stream(s)>>>((const __nv_bfloat16*)x, x_ld, n, h, w, c, k, pad, stride, ho, wo, (__nv_bfloat16*)out, m_pad);
  return check_launch();
}

int ppy_im2col_kmajor(const void* x, int x_ld, int n, int h, int w, int c, int k, int pad, void* out, long long m_pad, ppy_stream_t s) {
  PPY_REQUIRE(k >= 1 && 2 * pad == k - 1);            // the "same" stride-1 form: output grid = input grid
  return ppy_im2col_kmajor_strided(x, x_ld, n, h, w, c, k, 1, pad, out, m_pad, s);
}

// EMA of the trainable parameters (reference model/EMA.py:31-45), all tensors in one launch: shadow and parameters are
// addressed through a device table; the arithmetic keeps numpy's float32 order: f32(decay)*old + f32(1-decay)*new.
int ppy_ema_update(float* shadow_flat, const float* const* params, const long long* offsets, int num_tensors, float decay,
                   float one_minus_decay, ppy_stream_t s) {
  PPY_REQUIRE(shadow_flat && params && offsets && num_tensors > 0);
  ema_update_kernel<<<dim3(64, (unsigned)num_tensors), 256, 0, as_stream(s)>>>(shadow_flat, params, offsets, decay, one_minus_decay);
  return check_launch();
}

}  // extern "C"
