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
#include "nms_common.cuh"

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

constexpr int kDecPix = 16;          // pixels per chunk staged in shared memory
constexpr int kDecThreads = 256;

// objectness (IoU-aware fusion through the reference's clamp / de-sigmoid / re-sigmoid round trip) and box of one
// (pixel, anchor); t = the anchor's (tx, ty, tw, th, tobj), t_ioup = its IoU logit.  Returns the objectness.
__device__ __forceinline__ float decode_anchor(const DecodeArgs& p, const float* t, float t_ioup, int a, int gx, int gy, int img,
                                               float4* box) {
  const float sx = sigmoidf_ref(t[0]), sy = sigmoidf_ref(t[1]), ew = expf(t[2]), eh = expf(t[3]);
  float conf = sigmoidf_ref(t[4]);
  if (p.iou_aware) {                                                         // _de_sigmoid, head.py:97-109, then sigmoid again (:47)
    const float ip = sigmoidf_ref(t_ioup);
    const float eps = 1e-7f, inv_eps = 1.f / 1e-7f;
    float v = __fmul_rn(powf(conf, p.e_obj), powf(ip, p.e_iou));
    v = clampf_ref(v, eps, inv_eps);
    v = __fsub_rn(__fdiv_rn(1.f, v), 1.f);
    v = clampf_ref(v, eps, inv_eps);
    conf = sigmoidf_ref(-logf(v));
  }
  // (scale_x_y * sigmoid(t) + grid - (scale_x_y - 1) * 0.5) * stride      head.py:40
  float cx = __fmul_rn(__fsub_rn(__fadd_rn(__fmul_rn(p.sxy, sx), (float)gx), p.sxy_off), p.stride);
  float cy = __fmul_rn(__fsub_rn(__fadd_rn(__fmul_rn(p.sxy, sy), (float)gy), p.sxy_off), p.stride);
  float bw = __fmul_rn(ew, p.aw[a]);                                         // head.py:44
  float bh = __fmul_rn(eh, p.ah[a]);
  float hw = __fdiv_rn(bw, 2.f), hh = __fdiv_rn(bh, 2.f);
  float x0 = __fsub_rn(cx, hw), y0 = __fsub_rn(cy, hh), x1 = __fadd_rn(cx, hw), y1 = __fadd_rn(cy, hh);
  const float im_h = __ldg(p.im_size + 2 * img), im_w = __ldg(p.im_size + 2 * img + 1);
  const float fs = (float)p.size;
  x0 = __fmul_rn(__fdiv_rn(__fdiv_rn(x0, fs), p.stride), im_w);              // head.py:66-67
  y0 = __fmul_rn(__fdiv_rn(__fdiv_rn(y0, fs), p.stride), im_h);
  x1 = __fmul_rn(__fdiv_rn(__fdiv_rn(x1, fs), p.stride), im_w);
  y1 = __fmul_rn(__fdiv_rn(__fdiv_rn(y1, fs), p.stride), im_h);
  if (p.clip) {                                                              // head.py:73-76
    x0 = x0 < 0.f ? __fmul_rn(x0, 0.f) : x0;
    y0 = y0 < 0.f ? __fmul_rn(y0, 0.f) : y0;
    x1 = x1 > im_w ? im_w : x1;
    y1 = y1 > im_h ? im_h : y1;
  }
  *box = make_float4(x0, y0, x1, y1);
  return conf;
}

// Dense decode.  grid = (chunk walkers, images); a CTA stages kDecPix pixels (one contiguous run of the NHWC head
// output) in shared memory with 16-byte loads, one thread per (pixel, anchor) does the anchor-level math and writes the
// box, then all threads stream the kDecPix*A*C class scores (contiguous in the output) with one sigmoid each.  When a
// Matrix-NMS workspace is attached every score > score_threshold is also counted in the image's score histogram
// (shared-memory bins, flushed once per CTA), which saves the NMS front end one full pass over the scores.
// AN/NC > 0: compile-time anchor/class counts (3 x 80, the PP-YOLO head) so the index math has constant divisors.
template <int AN, int NC>
__global__ void __launch_bounds__(kDecThreads) yolo_decode_kernel(DecodeArgs p, CandSink sink) {
  extern __shared__ float dec_smem[];
  const int an = AN ? AN : p.an, nc = NC ? NC : p.nc;
  const int per = 5 + nc;
  const int first = p.iou_aware ? an : 0;
  const int hw = p.size * p.size;
  const int img = blockIdx.y, tid = threadIdx.x;
  float* stage = dec_smem;
  float* conf_s = stage + kDecPix * p.ld;
  unsigned int* hist_s = reinterpret_cast<unsigned int*>(conf_s + kDecPix * 4);
  const bool do_hist = sink.hist != nullptr;
  if (do_hist) for (int i = tid; i < kBins; i += kDecThreads) hist_s[i] = 0u;
  const bool vec = (p.ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.head) & 15) == 0;
  for (int chunk = blockIdx.x; chunk * kDecPix < hw; chunk += gridDim.x) {
    const int pix0 = chunk * kDecPix;
    const int npx = hw - pix0 < kDecPix ? hw - pix0 : kDecPix;
    const float* src = p.head + ((long long)img * hw + pix0) * p.ld;
    __syncthreads();                                        // previous chunk fully consumed (and the histogram zeroed)
    if (vec) {
      const float4* s4 = reinterpret_cast<const float4*>(src);
      for (int i = tid; i < npx * p.ld / 4; i += kDecThreads) reinterpret_cast<float4*>(stage)[i] = __ldg(s4 + i);
    } else {
      for (int i = tid; i < npx * p.ld; i += kDecThreads) stage[i] = __ldg(src + i);
    }
    __syncthreads();
    if (tid < npx * an) {
      const int pl = tid / an, a = tid - pl * an;
      const int pix = pix0 + pl;
      const float* px = stage + pl * p.ld;
      float4 box;
      const float conf = decode_anchor(p, px + first + a * per, p.iou_aware ? px[a] : 0.f, a, pix % p.size, pix / p.size, img, &box);
      conf_s[pl * 4 + a] = conf;
      reinterpret_cast<float4*>(p.boxes)[(long long)img * p.total_boxes + p.box_offset + (long long)pix * an + a] = box;
    }
    __syncthreads();
    const int tot = npx * an * nc;
    float* srow = p.scores + ((long long)img * p.total_boxes + p.box_offset + (long long)pix0 * an) * nc;
    for (int e = tid; e < tot; e += kDecThreads) {
      const int pl = e / (an * nc), r = e - pl * (an * nc);
      const int a = r / nc, c = r - a * nc;
      const float s = __fmul_rn(conf_s[pl * 4 + a], sigmoidf_ref(stage[pl * p.ld + first + a * per + 5 + c]));
      srow[e] = s;
      if (do_hist && s > sink.thr) atomicAdd(&hist_s[score_bin(s, sink.thr_bits, sink.shift)], 1u);
    }
  }
  if (do_hist) {
    __syncthreads();
    unsigned int* gh = sink.hist + (long long)img * kBins;
    for (int i = tid; i < kBins; i += kDecThreads)
      if (hist_s[i]) atomicAdd(&gh[i], hist_s[i]);
  }
}

// Two-kernel dense decode of the whole-network path (ppy_yolo_decode_hist).
//   yolo_anchor_kernel: one THREAD per (pixel, anchor): objectness + box with every lane busy (the transcendental-heavy
//                       part); writes the box and the objectness conf[img][box] (4 B per anchor of scratch).
//   yolo_scores_kernel: pure streaming pass, one thread per 4 consecutive classes of a box: 4 logits in, one 16-byte
//                       score store out, + the per-image score histogram in shared memory (flushed once per CTA).
// conf also lets the NMS front end skip whole boxes (score = conf * sigmoid(cls) <= conf).
__global__ void __launch_bounds__(256) yolo_anchor_kernel(DecodeArgs p, float* __restrict__ conf_out) {
  const long long total = (long long)p.n * p.size * p.size * p.an;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int per = 5 + p.nc;
  const int first = p.iou_aware ? p.an : 0;
  const long long pixg = gid / p.an;
  const int a = (int)(gid - pixg * p.an);
  const int img = (int)(pixg / (p.size * p.size)), pix = (int)(pixg - (long long)img * (p.size * p.size));
  const float* px = p.head + pixg * p.ld;
  const float* tp = px + first + a * per;
  const float t[5] = {__ldg(tp), __ldg(tp + 1), __ldg(tp + 2), __ldg(tp + 3), __ldg(tp + 4)};
  float4 box;
  const float conf = decode_anchor(p, t, p.iou_aware ? __ldg(px + a) : 0.f, a, pix % p.size, pix / p.size, img, &box);
  const long long row = (long long)img * p.total_boxes + p.box_offset + (long long)pix * p.an + a;
  reinterpret_cast<float4*>(p.boxes)[row] = box;
  conf_out[row] = conf;
}

template <int AN, int NC>
__global__ void __launch_bounds__(256) yolo_scores_kernel(DecodeArgs p, const float* __restrict__ conf_in, CandSink sink) {
  __shared__ unsigned int hist_s[kBins];
  const int an = AN ? AN : p.an, nc = NC ? NC : p.nc;
  const int per = 5 + nc, first = p.iou_aware ? an : 0, qpb = nc / 4;      // nc % 4 == 0 (checked by the host)
  const int img = blockIdx.y, tid = threadIdx.x;
  for (int i = tid; i < kBins; i += 256) hist_s[i] = 0u;
  __syncthreads();
  const int boxes_here = p.size * p.size * an;
  const int quads = boxes_here * qpb;
  const float* head = p.head + (long long)img * p.size * p.size * p.ld;
  const float* conf = conf_in + (long long)img * p.total_boxes + p.box_offset;
  float4* out = reinterpret_cast<float4*>(p.scores + ((long long)img * p.total_boxes + p.box_offset) * nc);
  for (int q = blockIdx.x * 256 + tid; q < quads; q += gridDim.x * 256) {
    const int b = q / qpb, c4 = (q - b * qpb) * 4;
    const int pix = b / an, a = b - pix * an;
    const float* lg = head + (long long)pix * p.ld + first + a * per + 5 + c4;
    const float l0 = __ldg(lg), l1 = __ldg(lg + 1), l2 = __ldg(lg + 2), l3 = __ldg(lg + 3);
    const float cf = __ldg(conf + b);
    float4 sc;
    sc.x = __fmul_rn(cf, sigmoidf_ref(l0)); sc.y = __fmul_rn(cf, sigmoidf_ref(l1));
    sc.z = __fmul_rn(cf, sigmoidf_ref(l2)); sc.w = __fmul_rn(cf, sigmoidf_ref(l3));
    out[q] = sc;
    if (cf > sink.thr) {                              // score <= conf: nothing to count otherwise
      if (sc.x > sink.thr) atomicAdd(&hist_s[score_bin(sc.x, sink.thr_bits, sink.shift)], 1u);
      if (sc.y > sink.thr) atomicAdd(&hist_s[score_bin(sc.y, sink.thr_bits, sink.shift)], 1u);
      if (sc.z > sink.thr) atomicAdd(&hist_s[score_bin(sc.z, sink.thr_bits, sink.shift)], 1u);
      if (sc.w > sink.thr) atomicAdd(&hist_s[score_bin(sc.w, sink.thr_bits, sink.shift)], 1u);
    }
  }
  __syncthreads();
  unsigned int* gh = sink.hist + (long long)img * kBins;
  for (int i = tid; i < kBins; i += 256)
    if (hist_s[i]) atomicAdd(&gh[i], hist_s[i]);
}

// Sparse variant for the whole-network path: one THREAD per (pixel, anchor) does the anchor-level math (IoU-aware
// objectness, box) with every lane busy, writes the box, and only anchors whose objectness exceeds the NMS score
// threshold get their C class scores evaluated (score = conf * sigmoid(cls) <= conf, so nothing is lost) -- by the
// whole warp, one class per lane.  Scores above the threshold go straight into the per-image Matrix-NMS candidate
// list + score histogram (nms_common.cuh); the dense [boxes x C] score tensor is never written and the class logits of
// background anchors are never read.  The arithmetic is the same sequence of fp32 operations as yolo_decode_kernel, so
// the candidates are bit-identical to thresholding the dense scores.
__global__ void __launch_bounds__(256) yolo_decode_sparse_kernel(DecodeArgs p, CandSink sink) {
  const int lane = threadIdx.x & 31;
  const long long total = (long long)p.n * p.size * p.size * p.an;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int per = 5 + p.nc;
  const int first = p.iou_aware ? p.an : 0;
  bool flag = false;
  float conf = 0.f;
  long long pixg = 0;
  int a = 0, img = 0, pix = 0;
  if (gid < total) {
    pixg = gid / p.an; a = (int)(gid - pixg * p.an);
    img = (int)(pixg / (p.size * p.size)); pix = (int)(pixg - (long long)img * (p.size * p.size));
    const float* px = p.head + pixg * p.ld;
    const float* tp = px + first + a * per;
    const float t[5] = {__ldg(tp), __ldg(tp + 1), __ldg(tp + 2), __ldg(tp + 3), __ldg(tp + 4)};
    float4 box;
    conf = decode_anchor(p, t, p.iou_aware ? __ldg(px + a) : 0.f, a, pix % p.size, pix / p.size, img, &box);
    reinterpret_cast<float4*>(p.boxes)[(long long)img * p.total_boxes + p.box_offset + (long long)pix * p.an + a] = box;
    flag = conf > sink.thr;
  }
  unsigned int todo = __ballot_sync(0xffffffffu, flag);
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const float conf_s = __shfl_sync(0xffffffffu, conf, src);
    const long long pixg_s = __shfl_sync(0xffffffffu, pixg, src);
    const int a_s = __shfl_sync(0xffffffffu, a, src), img_s = __shfl_sync(0xffffffffu, img, src);
    const int pix_s = __shfl_sync(0xffffffffu, pix, src);
    const float* cls = p.head + pixg_s * p.ld + first + a_s * per + 5;
    const unsigned int box_row = (unsigned int)(p.box_offset + pix_s * p.an + a_s);
    for (int c0 = 0; c0 < p.nc; c0 += 32) {
      const int c = c0 + lane;
      float s = 0.f;
      if (c < p.nc) s = __fmul_rn(conf_s, sigmoidf_ref(__ldg(cls + c)));
      const bool hit = c < p.nc && s > sink.thr;
      const unsigned int hits = __ballot_sync(0xffffffffu, hit);
      if (!hits) continue;
      unsigned int base = 0;
      if (lane == 0) base = atomicAdd(sink.count + img_s, (unsigned int)__popc(hits));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (hit) {
        const unsigned int slot = base + (unsigned int)__popc(hits & ((1u << lane) - 1u));
        if (slot < (unsigned int)sink.cap)
          sink.keys[(long long)img_s * sink.cap + slot] =
              ((unsigned long long)__float_as_uint(s) << 32) | (unsigned long long)(0xFFFFFFFFu - (box_row * (unsigned int)p.nc + (unsigned int)c));
        atomicAdd(sink.hist + (long long)img_s * kBins + score_bin(s, sink.thr_bits, sink.shift), 1u);
      }
    }
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

static int fill_decode_args(ppy::DecodeArgs& p, const float* head, int ld, int n, int size, int an_num, int num_classes,
                            const float* anchors, int stride, double scale_x_y, const float* im_size, int clip_bbox,
                            int iou_aware, double factor, float* boxes, float* scores, int box_offset, int total_boxes) {
  PPY_REQUIRE(head && anchors && im_size && boxes);
  PPY_REQUIRE(n > 0 && size > 0 && an_num > 0 && an_num <= 4 && num_classes > 0 && stride > 0);
  PPY_REQUIRE(ld >= an_num * (num_classes + (iou_aware ? 6 : 5)));
  PPY_REQUIRE(box_offset >= 0 && box_offset + size * size * an_num <= total_boxes);
  PPY_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15) == 0);
  PPY_REQUIRE(num_classes + 6 <= 96);
  p.head = head; p.ld = ld; p.n = n; p.size = size; p.an = an_num; p.nc = num_classes;
  for (int a = 0; a < 4; ++a) { p.aw[a] = a < an_num ? anchors[2 * a] : 0.f; p.ah[a] = a < an_num ? anchors[2 * a + 1] : 0.f; }
  p.stride = (float)stride; p.sxy = (float)scale_x_y;
  p.sxy_off = (float)((scale_x_y - 1.0) * 0.5);
  p.im_size = im_size; p.clip = clip_bbox; p.iou_aware = iou_aware;
  p.e_obj = (float)(1.0 - factor); p.e_iou = (float)factor;   // python-float exponents of head.py:125
  p.boxes = boxes; p.scores = scores; p.box_offset = box_offset; p.total_boxes = total_boxes;
  return PPY_OK;
}

static int launch_dense_decode(const ppy::DecodeArgs& p, const ppy::CandSink& sink, ppy_stream_t s) {
  using namespace ppy;
  const int hw = p.size * p.size;
  long long gx = ceil_div(148 * 6, p.n);                   // ~6 CTAs per SM over the whole batch
  const long long chunks = ceil_div(hw, kDecPix);
  if (gx > chunks) gx = chunks;
  if (gx < 1) gx = 1;
  const size_t smem = (size_t)(kDecPix * p.ld + kDecPix * 4) * 4 + (sink.hist ? kBins * 4 : 0);
  PPY_REQUIRE(smem <= 48 * 1024);
  dim3 grid((unsigned)gx, (unsigned)p.n);
  if (p.an == 3 && p.nc == 80) yolo_decode_kernel<3, 80><<<grid, kDecThreads, smem, as_stream(s)>>>(p, sink);
  else yolo_decode_kernel<0, 0><<<grid, kDecThreads, smem, as_stream(s)>>>(p, sink);
  return check_launch();
}

int ppy_yolo_decode(const float* head, int ld, int n, int size, int an_num, int num_classes, const float* anchors,
                    int stride, double scale_x_y, const float* im_size, int clip_bbox, int iou_aware, double factor,
                    float* boxes, float* scores, int box_offset, int total_boxes, ppy_stream_t s) {
  using namespace ppy;
  PPY_REQUIRE(scores);
  DecodeArgs p;
  int rc = fill_decode_args(p, head, ld, n, size, an_num, num_classes, anchors, stride, scale_x_y, im_size, clip_bbox,
                            iou_aware, factor, boxes, scores, box_offset, total_boxes);
  if (rc) return rc;
  CandSink none;
  memset(&none, 0, sizeof(none));
  return launch_dense_decode(p, none, s);
}

int ppy_yolo_decode_hist(const float* head, int ld, int n, int size, int an_num, int num_classes, const float* anchors,
                         int stride, double scale_x_y, const float* im_size, int clip_bbox, int iou_aware, double factor,
                         float* boxes, float* scores, int box_offset, int total_boxes, float score_threshold,
                         void* nms_workspace, ppy_stream_t s) {
  using namespace ppy;
  PPY_REQUIRE(scores && nms_workspace && score_threshold > 0.f);
  PPY_REQUIRE(num_classes % 4 == 0 && (reinterpret_cast<uintptr_t>(scores) & 15) == 0);
  DecodeArgs p;
  int rc = fill_decode_args(p, head, ld, n, size, an_num, num_classes, anchors, stride, scale_x_y, im_size, clip_bbox,
                            iou_aware, factor, boxes, scores, box_offset, total_boxes);
  if (rc) return rc;
  const CandSink sink = cand_carve(nms_workspace, n, kNmsKeyCap, score_threshold);   // the dense NMS workspace layout
  float* conf = reinterpret_cast<float*>(sink.keys + (size_t)n * kNmsKeyCap);         // conf[n][total_boxes] after the keys
  const long long total = (long long)n * size * size * an_num;
  yolo_anchor_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(s)>>>(p, conf);
  if ((rc = check_launch())) return rc;
  const long long quads = (long long)size * size * an_num * (num_classes / 4);
  long long gx = ceil_div(148 * 8, n);
  if (gx > ceil_div(quads, 256)) gx = ceil_div(quads, 256);
  dim3 grid((unsigned)gx, (unsigned)n);
  if (an_num == 3 && num_classes == 80) yolo_scores_kernel<3, 80><<<grid, 256, 0, as_stream(s)>>>(p, conf, sink);
  else yolo_scores_kernel<0, 0><<<grid, 256, 0, as_stream(s)>>>(p, conf, sink);
  return check_launch();
}

int ppy_nms_candidate_workspace_bytes(int n, int cap, size_t* bytes) {
  PPY_REQUIRE(bytes && n > 0 && cap > 0);
  *bytes = ppy::cand_workspace_bytes(n, cap);
  return PPY_OK;
}

int ppy_nms_candidates_reset(void* workspace, int n, int cap, ppy_stream_t s) {
  using namespace ppy;
  PPY_REQUIRE(workspace && n > 0 && cap > 0);
  // histogram + counters live at the head of the workspace, contiguous
  return check_cuda(cudaMemsetAsync(workspace, 0, sizeof(unsigned int) * (size_t)n * kBins + cand_count_bytes(n), as_stream(s)));
}

int ppy_yolo_decode_candidates(const float* head, int ld, int n, int size, int an_num, int num_classes,
                               const float* anchors, int stride, double scale_x_y, const float* im_size, int clip_bbox,
                               int iou_aware, double factor, float* boxes, int box_offset, int total_boxes,
                               float score_threshold, void* workspace, int cap, ppy_stream_t s) {
  using namespace ppy;
  PPY_REQUIRE(workspace && cap > 0 && score_threshold > 0.f);
  PPY_REQUIRE((long long)total_boxes * num_classes < 0xFFFFFFFFll);
  DecodeArgs p;
  int rc = fill_decode_args(p, head, ld, n, size, an_num, num_classes, anchors, stride, scale_x_y, im_size, clip_bbox,
                            iou_aware, factor, boxes, nullptr, box_offset, total_boxes);
  if (rc) return rc;
  const CandSink sink = cand_carve(workspace, n, cap, score_threshold);
  const long long total = (long long)n * size * size * an_num;
  yolo_decode_sparse_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(s)>>>(p, sink);
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
