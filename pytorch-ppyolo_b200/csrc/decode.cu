// Head post-processing for sm_100a: IoU-aware objectness fusion + YOLO box decode + score product in one
// pass over the NHWC fp32 head output (reference model/head.py:21-141).  HBM-bound: reads
// (A*(6+C))*4 B per pixel once, writes boxes (16 B) and C scores per (pixel, anchor); one warp per
// (pixel, anchor) so the C-wide score row is a single coalesced store.
//
// The arithmetic mirrors the reference's fp32 operation order (explicit _rn intrinsics keep nvcc from
// contracting mul+add into FMA), so results differ from the CPU path only by the libm ulp of
// exp/log/pow.
#include <math_constants.h>
#include "common.cuh"

namespace ppy {
namespace {

// 1/(1+e^-x); __frcp_rn is the correctly rounded reciprocal, i.e. bit-identical to __fdiv_rn(1, .) and cheaper
__device__ __forceinline__ float sigmoidf_ref(float x) { return __frcp_rn(__fadd_rn(1.f, expf(-x))); }

__device__ __forceinline__ float clampf_ref(float v, float lo, float hi) {  // torch.clamp, NaN propagates
  if (v != v) return v;
  return v < lo ? lo : (v > hi ? hi : v);
}

// _de_sigmoid (head.py:97-109) of obj^(1-f) * ioup^f (head.py:125)
__device__ __forceinline__ float fused_obj_logit(float t_obj, float t_ioup, float e_obj, float e_iou) {
  float obj = sigmoidf_ref(t_obj);
  float ip = sigmoidf_ref(t_ioup);
  float v = __fmul_rn(powf(obj, e_obj), powf(ip, e_iou));
  const float eps = 1e-7f, inv_eps = 1.f / 1e-7f;
  v = clampf_ref(v, eps, inv_eps);
  v = __fsub_rn(__fdiv_rn(1.f, v), 1.f);
  v = clampf_ref(v, eps, inv_eps);
  return -logf(v);
}

struct DecodeArgs {
  const float* head; int ld; int n; int size; int an; int nc;
  float aw[4]; float ah[4];
  float stride; float sxy; float sxy_off;  // (sxy - 1) * 0.5 rounded to fp32 like the reference's python float
  const float* im_size; int clip; int iou_aware; float e_obj; float e_iou;
  float* boxes; float* scores; int box_offset; int total_boxes;
};

constexpr int kDecodeWarps = 8;
constexpr int kMaxPixelFloats = 4 * 96;    // an <= 4, nc + 6 <= 96 channels per anchor staged per warp

// One warp per pixel: the pixel's A*(5|6+C) logits are staged once in shared memory with coalesced loads
// (all in flight together), every lane then sigmoids its share of the A*C class logits; the A score rows of a
// pixel are contiguous in the output (box index = pixel*A + a), so the warp writes one contiguous run.
__global__ void __launch_bounds__(kDecodeWarps * 32) yolo_decode_kernel(DecodeArgs p) {
  __shared__ float stage[kDecodeWarps][kMaxPixelFloats];
  __shared__ float conf_s[kDecodeWarps][4];
  __shared__ float tmp_s[kDecodeWarps][32];      // [0,24): step-1 results (a*6+f), [24,32): step-2 powers
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long npix = (long long)p.n * p.size * p.size;
  const int per = 5 + p.nc;
  const int chans = p.an * (per + (p.iou_aware ? 1 : 0));
  const int first = p.iou_aware ? p.an : 0;
  float* st = stage[wib];
  for (long long pixg = (long long)blockIdx.x * kDecodeWarps + wib; pixg < npix; pixg += (long long)gridDim.x * kDecodeWarps) {
    const int img = (int)(pixg / (p.size * p.size));
    const int pix = (int)(pixg % (p.size * p.size));
    const int gx = pix % p.size, gy = pix / p.size;
    const float* px = p.head + pixg * p.ld;
    for (int i = lane; i < chans; i += 32) st[i] = __ldg(px + i);
    __syncwarp();
    // anchor-level math spread over lanes in lock-step (same instruction stream for every active lane):
    //   step 1, lanes 0..6A-1: item (a, f) -> f<2: sigmoid(t_xy), f in {2,3}: exp(t_wh), f=4: sigmoid(obj), f=5: sigmoid(ioup)
    //   step 2, lanes 0..2A-1: pow(sigmoid(obj), 1-f) and pow(sigmoid(ioup), f)            (IoU-aware only)
    //   step 3, lanes 0..A-1 : clamp / de-sigmoid / re-sigmoid of the fused objectness, box assembly and clip
    if (lane < 6 * p.an) {
      const int a = lane / 6, f = lane - 6 * a;
      const float t = (f == 5) ? (p.iou_aware ? st[a] : 0.f) : st[first + a * per + f];
      const bool is_exp = (f == 2 || f == 3);
      const float e = expf(is_exp ? t : -t);
      tmp_s[wib][lane] = is_exp ? e : __frcp_rn(__fadd_rn(1.f, e));
    }
    __syncwarp();
    if (p.iou_aware && lane < 2 * p.an) {
      const int a = lane >> 1, j = lane & 1;
      tmp_s[wib][24 + lane] = powf(tmp_s[wib][6 * a + 4 + j], j ? p.e_iou : p.e_obj);
    }
    __syncwarp();
    if (lane < p.an) {
      const int a = lane;
      const float* r = tmp_s[wib] + 6 * a;
      float conf = r[4];
      if (p.iou_aware) {                                                       // _de_sigmoid, head.py:97-109, then sigmoid again (:47)
        const float eps = 1e-7f, inv_eps = 1.f / 1e-7f;
        float v = __fmul_rn(tmp_s[wib][24 + 2 * a], tmp_s[wib][24 + 2 * a + 1]);
        v = clampf_ref(v, eps, inv_eps);
        v = __fsub_rn(__fdiv_rn(1.f, v), 1.f);
        v = clampf_ref(v, eps, inv_eps);
        conf = sigmoidf_ref(-logf(v));
      }
      conf_s[wib][a] = conf;
      // (scale_x_y * sigmoid(t) + grid - (scale_x_y - 1) * 0.5) * stride      head.py:40
      float cx = __fmul_rn(__fsub_rn(__fadd_rn(__fmul_rn(p.sxy, r[0]), (float)gx), p.sxy_off), p.stride);
      float cy = __fmul_rn(__fsub_rn(__fadd_rn(__fmul_rn(p.sxy, r[1]), (float)gy), p.sxy_off), p.stride);
      float bw = __fmul_rn(r[2], p.aw[a]);                                     // head.py:44
      float bh = __fmul_rn(r[3], p.ah[a]);
      float hw = __fdiv_rn(bw, 2.f), hh = __fdiv_rn(bh, 2.f);
      float x0 = __fsub_rn(cx, hw), y0 = __fsub_rn(cy, hh), x1 = __fadd_rn(cx, hw), y1 = __fadd_rn(cy, hh);
      const float im_h = __ldg(p.im_size + 2 * img), im_w = __ldg(p.im_size + 2 * img + 1);
      const float fs = (float)p.size;
      x0 = __fmul_rn(__fdiv_rn(__fdiv_rn(x0, fs), p.stride), im_w);            // head.py:66-67
      y0 = __fmul_rn(__fdiv_rn(__fdiv_rn(y0, fs), p.stride), im_h);
      x1 = __fmul_rn(__fdiv_rn(__fdiv_rn(x1, fs), p.stride), im_w);
      y1 = __fmul_rn(__fdiv_rn(__fdiv_rn(y1, fs), p.stride), im_h);
      if (p.clip) {                                                            // head.py:73-76
        x0 = x0 < 0.f ? __fmul_rn(x0, 0.f) : x0;
        y0 = y0 < 0.f ? __fmul_rn(y0, 0.f) : y0;
        x1 = x1 > im_w ? im_w : x1;
        y1 = y1 > im_h ? im_h : y1;
      }
      const long long row = (long long)img * p.total_boxes + p.box_offset + (long long)pix * p.an + a;
      reinterpret_cast<float4*>(p.boxes)[row] = make_float4(x0, y0, x1, y1);
    }
    __syncwarp();
    float* srow = p.scores + ((long long)img * p.total_boxes + p.box_offset + (long long)pix * p.an) * p.nc;
    const int total = p.an * p.nc;
    for (int i = lane; i < total; i += 32) {
      const int a = i / p.nc, c = i - a * p.nc;
      srow[i] = __fmul_rn(conf_s[wib][a], sigmoidf_ref(st[first + a * per + 5 + c]));
    }
    __syncwarp();
  }
}

__global__ void iou_aware_kernel(const float* __restrict__ x, int x_ld, float* __restrict__ y, int y_ld, long long pixels,
                                 int an, int nc, float e_obj, float e_iou) {
  const int per = 5 + nc;
  const long long total = pixels * an * per;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long pix = i / (an * per);
    int rem = (int)(i % (an * per));
    int a = rem / per, f = rem % per;
    const float* px = x + pix * x_ld;
    float v = __ldg(px + an + rem);
    if (f == 4) v = fused_obj_logit(v, __ldg(px + a), e_obj, e_iou);
    y[pix * y_ld + rem] = v;
  }
}

}  // namespace
}  // namespace ppy

extern "C" {

int ppy_yolo_decode(const float* head, int ld, int n, int size, int an_num, int num_classes, const float* anchors,
                    int stride, double scale_x_y, const float* im_size, int clip_bbox, int iou_aware, double factor,
                    float* boxes, float* scores, int box_offset, int total_boxes, ppy_stream_t s) {
  using namespace ppy;
  PPY_REQUIRE(head && anchors && im_size && boxes && scores);
  PPY_REQUIRE(n > 0 && size > 0 && an_num > 0 && an_num <= 4 && num_classes > 0 && stride > 0);
  PPY_REQUIRE(ld >= an_num * (num_classes + (iou_aware ? 6 : 5)));
  PPY_REQUIRE(box_offset >= 0 && box_offset + size * size * an_num <= total_boxes);
  PPY_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15) == 0);
  DecodeArgs p;
  p.head = head; p.ld = ld; p.n = n; p.size = size; p.an = an_num; p.nc = num_classes;
  for (int a = 0; a < 4; ++a) { p.aw[a] = a < an_num ? anchors[2 * a] : 0.f; p.ah[a] = a < an_num ? anchors[2 * a + 1] : 0.f; }
  p.stride = (float)stride; p.sxy = (float)scale_x_y;
  p.sxy_off = (float)((scale_x_y - 1.0) * 0.5);
  p.im_size = im_size; p.clip = clip_bbox; p.iou_aware = iou_aware;
  p.e_obj = (float)(1.0 - factor); p.e_iou = (float)factor;   // python-float exponents of head.py:125
  p.boxes = boxes; p.scores = scores; p.box_offset = box_offset; p.total_boxes = total_boxes;
  PPY_REQUIRE(num_classes + 6 <= 96);
  long long pixels = (long long)n * size * size;
  long long blocks = ceil_div(pixels, kDecodeWarps);
  if (blocks > 148 * 8) blocks = 148 * 8;
  yolo_decode_kernel<<<(unsigned)blocks, kDecodeWarps * 32, 0, as_stream(s)>>>(p);
  return check_launch();
}

int ppy_iou_aware_score(const float* x, int x_ld, float* y, int y_ld, long long pixels, int an_num, int num_classes,
                        double factor, ppy_stream_t s) {
  using namespace ppy;
  PPY_REQUIRE(x && y && pixels > 0 && an_num > 0 && num_classes > 0);
  PPY_REQUIRE(x_ld >= an_num * (num_classes + 6) && y_ld >= an_num * (num_classes + 5));
  long long total = pixels * an_num * (5 + num_classes);
  long long blocks = ceil_div(total, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  iou_aware_kernel<<<(unsigned)blocks, 256, 0, as_stream(s)>>>(x, x_ld, y, y_ld, pixels, an_num, num_classes,
                                                              (float)(1.0 - factor), (float)factor);
  return check_launch();
}

}  // extern "C"
