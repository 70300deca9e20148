// First layer of the ResNet-vd deep stem fused with the input layout change: reads the NCHW fp32 image batch
// exactly as Decode.predict uploads it (reference model/decode_np.py:142-147), applies conv 3x3 / stride 2 / pad 1
// (3 -> 32, reference model/resnet_vd.py:100) + folded BN + activation in fp32 SIMT math and writes NHWC
// (bf16 or fp32).  K = 27 is far too small for the tensor pipe and the layer is HBM-bound (142 MB in, 189 MB out
// at bs=32 x 608^2); the weights ride in the kernel parameter (constant bank), so every FFMA takes its weight as a
// constant operand and each thread keeps its 32 output channels of one pixel in registers.
#include "common.cuh"

namespace ppy {
namespace {

constexpr int STEM_CIN = 3, STEM_COUT = 32, STEM_TAPS = 27;

struct StemParams {
  float w[STEM_TAPS][STEM_COUT];   // [c*9 + ky*3 + kx][cout]
  float scale[STEM_COUT];
  float shift[STEM_COUT];
};

template <typename T>
__global__ void __launch_bounds__(128) stem_conv_kernel(const float* __restrict__ x, int n, int h, int w, int ho, int wo,
                                                        const __grid_constant__ StemParams prm, float slope, T* __restrict__ y,
                                                        int y_ld) {
  const long long total = (long long)n * ho * wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % wo), oy = (int)((i / wo) % ho), img = (int)(i / ((long long)wo * ho));
    float acc[STEM_COUT];
#pragma unroll
    for (int co = 0; co < STEM_COUT; ++co) acc[co] = 0.f;
    const float* xi = x + (long long)img * STEM_CIN * h * w;
#pragma unroll
    for (int c = 0; c < STEM_CIN; ++c) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * 2 - 1 + ky;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = ox * 2 - 1 + kx;
          float v = 0.f;
          if (iy >= 0 && iy < h && ix >= 0 && ix < w) v = __ldg(xi + ((long long)c * h + iy) * w + ix);
#pragma unroll
          for (int co = 0; co < STEM_COUT; ++co) acc[co] = fmaf(v, prm.w[c * 9 + ky * 3 + kx][co], acc[co]);
        }
      }
    }
    T* dst = y + i * y_ld;
#pragma unroll
    for (int co = 0; co < STEM_COUT; co += 16 / sizeof(T)) {
      uint4 raw;
      T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
      for (int k = 0; k < (int)(16 / sizeof(T)); ++k) {
        float f = acc[co + k] * prm.scale[co + k] + prm.shift[co + k];
        f = f > 0.f ? f : f * slope;
        e[k] = from_f<T>(f);
      }
      *reinterpret_cast<uint4*>(dst + co) = raw;
    }
  }
}

}  // namespace
}  // namespace ppy

extern "C" int ppy_stem_conv3x3s2(const float* x_nchw, int n, int h, int w, const float* weight_oihw_host,
                                  const float* scale_host, const float* shift_host, int cout, int act, void* y, int y_ld,
                                  int y_dtype, ppy_stream_t s) {
  using namespace ppy;
  PPY_REQUIRE(x_nchw && weight_oihw_host && scale_host && shift_host && y);
  PPY_REQUIRE(n > 0 && h > 1 && w > 1 && cout == STEM_COUT && y_ld >= cout);
  PPY_REQUIRE(act == PPY_ACT_NONE || act == PPY_ACT_RELU || act == PPY_ACT_LEAKY);
  PPY_REQUIRE((reinterpret_cast<uintptr_t>(y) & 15) == 0 && (y_ld * dtype_size(y_dtype)) % 16 == 0);
  StemParams prm;
  for (int co = 0; co < STEM_COUT; ++co) {
    for (int t = 0; t < STEM_TAPS; ++t) prm.w[t][co] = weight_oihw_host[co * STEM_TAPS + t];   // OIHW: [co][c][ky][kx]
    prm.scale[co] = scale_host[co];
    prm.shift[co] = shift_host[co];
  }
  const float slope = act == PPY_ACT_RELU ? 0.f : (act == PPY_ACT_LEAKY ? 0.1f : 1.f);
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  const long long total = (long long)n * ho * wo;
  long long blocks = ceil_div(total, 128);
  if (blocks > 148 * 64) blocks = 148 * 64;
  if (y_dtype == PPY_BF16)
    stem_conv_kernel<__nv_bfloat16><<<(unsigned)blocks, 128, 0, as_stream(s)>>>(x_nchw, n, h, w, ho, wo, prm, slope,
                                                                                (__nv_bfloat16*)y, y_ld);
  else if (y_dtype == PPY_F32)
    stem_conv_kernel<float><<<(unsigned)blocks, 128, 0, as_stream(s)>>>(x_nchw, n, h, w, ho, wo, prm, slope, (float*)y, y_ld);
  else return PPY_ERR_INVALID;
  return check_launch();
}
