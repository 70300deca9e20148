// DCNv2 backward, sampling side (training with an unfrozen stage 5; reference model/custom_layers.py:551-677 differentiated by
// torch autograd there -- gather_nd's index backward, the bilinear weights, the clamp and the sigmoid).
//
// The layer is  y[m][o] = sum_{t,c} col[m][t][c] * W[o][c][t],  col[m][t][c] = mask[m][t] * sum_q wq[m][t][q] * x[corner_q(m,t)][c].
// Its two GEMM-shaped gradients run on the tcgen05 conv kernel (conv_autograd.py):
//     dW   = dY^T [O x M] . col [M x 9C]       (col rebuilt by ppy_dcn_gather, K = the pixels, partial-sum launches)
//     dcol = dY   [M x O] . Wt  [O x 9C]       (a 1x1 conv with 9C output channels, fp32 result)
// and this kernel turns dcol into the gradients of everything the sampler read -- one warp per (output pixel, tap):
//     s_q      = <dcol[m][t][:], x[corner_q][:]>                      four dot products over the channels (shuffle reduction)
//     d mask   = sum_q wq * s_q ;  d logit = d mask * mask * (1 - mask)
//     d pos_y  = mask * (hx * (s2 - s0) + lx * (s3 - s1))            (ly = frac(pos_y): w0 = hy*hx, w1 = hy*lx, w2 = ly*hx, w3 = ly*lx)
//     d pos_x  = mask * (hy * (s1 - s0) + ly * (s3 - s2))
//     dx[corner_q][:] += mask * wq * dcol[m][t][:]                    16-byte vector reductions (red.global.add.v4.f32) into fp32 NHWC
// Corners outside the image contribute nothing and receive nothing; a sample the reference clamps (pos outside
// [-pad, H + pad - 1], :571-574) has all four corners in the zero border, so its offset gradient is zero here as it is there.
// Same position arithmetic as the forward kernels (layout_pool.cu dcn_gather_kernel, dcn_umma.cu).
#include "common.cuh"

namespace ppy {
namespace {

template <typename T> struct VecOf;
template <> struct VecOf<float> { static constexpr int N = 4; };
template <> struct VecOf<__nv_bfloat16> { static constexpr int N = 8; };

template <typename T>
__device__ __forceinline__ void load_row_vec(const T* p, float* out);
template <>
__device__ __forceinline__ void load_row_vec<float>(const float* p, float* out) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
template <>
__device__ __forceinline__ void load_row_vec<__nv_bfloat16>(const __nv_bfloat16* p, float* out) {
  const uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    out[2 * i] = f.x; out[2 * i + 1] = f.y;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) dcn_backward_sample_kernel(const T* __restrict__ x, int x_ld, int n, int h, int w, int c,
                                                                  const float* __restrict__ om, int om_ld, int k, int stride, int pad,
                                                                  int ho, int wo, const float* __restrict__ dcol,
                                                                  float* __restrict__ dx, int dx_ld, float* __restrict__ dom) {
  constexpr int V = VecOf<T>::N;
  const int lane = threadIdx.x & 31;
  const int taps = k * k;
  const long long total = (long long)n * ho * wo * taps;
  const int cv = c / V;
  for (long long wi = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wi < total; wi += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int tap = (int)(wi % taps);
    const long long m = wi / taps;
    const int ox = (int)(m % wo), oy = (int)((m / wo) % ho), img = (int)(m / ((long long)wo * ho));
    const float* o = om + m * om_ld;
    const float dy = __ldg(o + 2 * tap), dxo = __ldg(o + 2 * tap + 1), ml = __ldg(o + 2 * taps + tap);
    const float mask = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-ml)));
    const float py = (float)(oy * stride - pad + tap / k) + dy, px = (float)(ox * stride - pad + tap % k) + dxo;
    const float fy = floorf(py), fx = floorf(px);
    const float ly = py - fy, lx = px - fx, hy = 1.f - ly, hx = 1.f - lx;
    const int y0 = (int)fminf(fmaxf(fy, -2.f), (float)h), x0 = (int)fminf(fmaxf(fx, -2.f), (float)w);
    const float wq[4] = {hy * hx, hy * lx, ly * hx, ly * lx};
    long long corner[4];
    bool ok[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int yy = y0 + (q >> 1), xx = x0 + (q & 1);
      ok[q] = yy >= 0 && yy < h && xx >= 0 && xx < w;
      corner[q] = ((long long)img * h + min(max(yy, 0), h - 1)) * w + min(max(xx, 0), w - 1);
    }
    const float* g = dcol + (m * taps + tap) * (long long)c;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int v = lane; v < cv; v += 32) {
      float gv[V];
#pragma unroll
      for (int e = 0; e < V; e += 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(g + v * V + e));
        gv[e] = t.x; gv[e + 1] = t.y; gv[e + 2] = t.z; gv[e + 3] = t.w;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (!ok[q]) continue;
        float xv[V];
        load_row_vec<T>(x + corner[q] * x_ld + v * V, xv);
        float acc = 0.f;
#pragma unroll
        for (int e = 0; e < V; ++e) acc = fmaf(gv[e], xv[e], acc);
        s[q] += acc;
        const float wm = wq[q] * mask;
        if (wm != 0.f) {
          float* d = dx + corner[q] * dx_ld + v * V;
#pragma unroll
          for (int e = 0; e < V; e += 4)
            atomicAdd(reinterpret_cast<float4*>(d + e), make_float4(wm * gv[e], wm * gv[e + 1], wm * gv[e + 2], wm * gv[e + 3]));
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) s[q] += __shfl_xor_sync(0xffffffffu, s[q], sh);
    if (lane == 0) {
      float* d = dom + m * om_ld;
      const float dmask = wq[0] * s[0] + wq[1] * s[1] + wq[2] * s[2] + wq[3] * s[3];
      d[2 * tap] = mask * (hx * (s[2] - s[0]) + lx * (s[3] - s[1]));
      d[2 * tap + 1] = mask * (hy * (s[1] - s[0]) + ly * (s[3] - s[2]));
      d[2 * taps + tap] = dmask * mask * (1.f - mask);
    }
  }
}

}  // namespace
}  // namespace ppy

extern "C" int ppy_dcn_backward_sample(const void* x, int x_ld, int n, int h, int w, int c, const float* offset_mask, int om_ld, int k,
                                       int stride, int pad, const float* dcol, float* dx, int dx_ld, float* d_offset_mask, int dtype,
                                       ppy_stream_t s) {
  using namespace ppy;
  PPY_REQUIRE(x && offset_mask && dcol && dx && d_offset_mask && n > 0 && h > 0 && w > 0 && c > 0 && k > 0 && stride > 0 && pad >= 0);
  PPY_REQUIRE(dtype == PPY_F32 || dtype == PPY_BF16);
  const int vec = dtype == PPY_BF16 ? 8 : 4;
  PPY_REQUIRE(c % vec == 0 && x_ld >= c && (x_ld * dtype_size(dtype)) % 16 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0);
  PPY_REQUIRE(dx_ld >= c && dx_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(dx) & 15) == 0 && (reinterpret_cast<uintptr_t>(dcol) & 15) == 0);
  PPY_REQUIRE(om_ld >= 3 * k * k);
  const int ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;
  PPY_REQUIRE(ho > 0 && wo > 0);
  const long long warps = (long long)n * ho * wo * k * k;
  long long blocks = ceil_div(warps, 8);
  if (blocks > 148 * 64) blocks = 148 * 64;
  if (dtype == PPY_BF16)
    dcn_backward_sample_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, as_stream(s)>>>(
        (const __nv_bfloat16*)x, x_ld, n, h, w, c, offset_mask, om_ld, k, stride, pad, ho, wo, dcol, dx, dx_ld, d_offset_mask);
  else
    dcn_backward_sample_kernel<float><<<(unsigned)blocks, 256, 0, as_stream(s)>>>((const float*)x, x_ld, n, h, w, c, offset_mask, om_ld, k,
                                                                                 stride, pad, ho, wo, dcol, dx, dx_ld, d_offset_mask);
  return check_launch();
}
