// Image pre-processing on the GPU: the reference's ResizeImage step -- cv2.resize of the decoded uint8 image to
// target_size x target_size with interp = cv2.INTER_CUBIC (tools/transform.py:923-1018, model/decode_np.py:125-134) -- for a
// whole batch of differently sized images in one launch, fused with decodeImage's BGR -> RGB swap.  The output is the uint8 HWC
// batch the fused stem kernel (ppy_stem_conv3x3s2_u8) normalises and permutes, so the host only uploads the ORIGINAL images.
//
// Arithmetic = OpenCV's own bicubic resize for 8-bit images (imgproc/src/resize.cpp, resizeGeneric_ + HResizeCubic +
// VResizeCubicVec_32s8u), restated operation by operation:
//   * per axis: f = (float)((d + 0.5) * scale - 0.5), s = floor(f), f -= s; Catmull-Rom-like weights with A = -0.75 in float,
//     in OpenCV's operation order, times 2048 and rounded to nearest-even into 16-bit fixed point; taps s-1 .. s+2 clamped to the
//     image (replicated border);
//   * horizontal pass in 32-bit integers (sum of pixel * weight);
//   * vertical pass in float: S0*b0 + (S1*b1 + (S2*b2 + S3*b3)) with b = weight / 2^22, every product and sum rounded separately
//     (no fused multiply-add: OpenCV's baseline code path), round to nearest-even, saturate to [0, 255].
// This equals cv2.resize bit for bit when OpenCV runs its own code (cv2.ipp.setUseIPP(False), or any build without Intel IPP);
// the pip wheels route 8-bit bicubic through Intel's closed-source IPP primitive, which differs from OpenCV's own code by +-1 on
// about 3 % of the pixels -- the same tolerance applies to this kernel (tests/test_gpu_preprocess.py measures both).
#include "common.cuh"

namespace ppy {
namespace {

__device__ __forceinline__ void cubic_weights(float x, int (&w)[4]) {
  const float A = -0.75f;
  const float x1 = __fadd_rn(x, 1.f);
  float c[4];
  // ((A*(x + 1) - 5*A)*(x + 1) + 8*A)*(x + 1) - 4*A
  c[0] = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, x1), 5.f * A), x1), 8.f * A), x1), 4.f * A);
  // ((A + 2)*x - (A + 3))*x*x + 1
  c[1] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.f, x), A + 3.f), x), x), 1.f);
  // ((A + 2)*(1 - x) - (A + 3))*(1 - x)*(1 - x) + 1
  const float y = __fsub_rn(1.f, x);
  c[2] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.f, y), A + 3.f), y), y), 1.f);
  c[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.f, c[0]), c[1]), c[2]);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int v = __float2int_rn(__fmul_rn(c[k], 2048.f));      // saturate_cast<short>(cvRound(.))
    w[k] = v < -32768 ? -32768 : (v > 32767 ? 32767 : v);
  }
}

__device__ __forceinline__ void axis_taps(int d, double scale, int ssize, int (&idx)[4], int (&w)[4]) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  const int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  cubic_weights(f, w);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int t = s - 1 + k;
    idx[k] = t < 0 ? 0 : (t >= ssize ? ssize - 1 : t);
  }
}

// meta[i] = {byte offset of image i in `src`, height, width}; images are tightly packed HWC uint8 with 3 channels.
// grid.z = image, one thread per output pixel.
__global__ void __launch_bounds__(256) resize_cubic_u8_kernel(const uint8_t* __restrict__ src, const long long* __restrict__ meta,
                                                              uint8_t* __restrict__ dst, int dsize, int swap_rb) {
  const int img = blockIdx.z;
  const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
  if (dx >= dsize) return;
  const long long off = meta[3 * img];
  const int sh = (int)meta[3 * img + 1], sw = (int)meta[3 * img + 2];
  const uint8_t* s = src + off;
  // cv2.resize(img, None, None, fx = S / w, fy = S / h): scale = 1 / fx in double
  const double scale_x = 1.0 / ((double)dsize / (double)sw), scale_y = 1.0 / ((double)dsize / (double)sh);
  int xi[4], xw[4], yi[4], yw[4];
  axis_taps(dx, scale_x, sw, xi, xw);
  axis_taps(dy, scale_y, sh, yi, yw);
  float acc[3] = {0.f, 0.f, 0.f};
  const float inv = 1.f / 4194304.f;
#pragma unroll
  for (int r = 3; r >= 0; --r) {
    const uint8_t* row = s + (size_t)yi[r] * sw * 3;
    int h0 = 0, h1 = 0, h2 = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint8_t* p = row + xi[k] * 3;
      h0 += (int)p[0] * xw[k]; h1 += (int)p[1] * xw[k]; h2 += (int)p[2] * xw[k];
    }
    const float b = __fmul_rn((float)yw[r], inv);
    if (r == 3) {
      acc[0] = __fmul_rn((float)h0, b); acc[1] = __fmul_rn((float)h1, b); acc[2] = __fmul_rn((float)h2, b);
    } else {
      acc[0] = __fadd_rn(__fmul_rn((float)h0, b), acc[0]);
      acc[1] = __fadd_rn(__fmul_rn((float)h1, b), acc[1]);
      acc[2] = __fadd_rn(__fmul_rn((float)h2, b), acc[2]);
    }
  }
  uint8_t o[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    int v = __float2int_rn(acc[c]);
    v = v < -32768 ? -32768 : (v > 32767 ? 32767 : v);        // v_pack: saturate to int16, then to uint8
    o[c] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
  }
  uint8_t* d = dst + (((size_t)img * dsize + dy) * dsize + dx) * 3;
  if (swap_rb) { d[0] = o[2]; d[1] = o[1]; d[2] = o[0]; }
  else { d[0] = o[0]; d[1] = o[1]; d[2] = o[2]; }
}

}  // namespace
}  // namespace ppy

extern "C" {
using namespace ppy;

int ppy_resize_cubic_u8_batch(const uint8_t* src, const long long* meta, int n, uint8_t* dst, int dsize, int swap_rb, ppy_stream_t s) {
  PPY_REQUIRE(src && meta && dst && n > 0 && n <= 65535 && dsize > 0 && dsize <= 65535);
  dim3 grid((unsigned)ceil_div(dsize, 256), (unsigned)dsize, (unsigned)n);
  resize_cubic_u8_kernel<<<grid, 256, 0, as_stream(s)>>>(src, meta, dst, dsize, swap_rb);
  return check_launch();
}

}  // extern "C"
