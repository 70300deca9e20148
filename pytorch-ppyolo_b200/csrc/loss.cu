// Fused fine-grained YOLOv3 loss of PP-YOLO training (reference model/losses.py:121-356 `_get_fine_grained_loss` +
// `_calc_obj_loss`, model/iou_losses.py IouLoss :15-191 / IouAwareLoss :194-246) for ONE output scale: one forward kernel (six
// loss sums + the no-object mask) and one backward kernel (gradient w.r.t. the raw head output, every channel written once),
// replacing ~150 small tensor kernels per scale.  Plus the target assignment of the data pipeline (tools/transform.py:1318-1421,
// Gt2YoloTargetSingle) as a kernel.
//
// Layouts (the reference's): out [N, A*(5|6+C), S, S] fp32 NCHW -- with iou_aware the first A channels are the IoU logits, then
// per anchor (x, y, w, h, obj, cls[C]); target [N, A, 6+C, S, S] = (tx, ty, tw, th, tscale, tobj, one-hot class);
// gt_box [N, G, 4] normalised (cx, cy, w, h).  One warp owns one (image, anchor, row): the IoU-aware loss is summed over the
// LAST axis before it meets the objectness target (model/iou_losses.py:241-243), which couples the cells of a row through
// T_row = sum_w tobj -- a warp reduction here.  All arithmetic is fp32 in the reference's operation order; the loss sums are
// accumulated in fp64 and reduced in a fixed order (per-block partials, last block adds them up): run-to-run deterministic.
#include <math.h>
#include "common.cuh"

namespace ppy {
namespace {

constexpr int MAX_ANCHORS = 8;
constexpr int WARPS = 8;
constexpr int MAX_GT = 128;

struct LossCfg {
  int n, a, c, s, g, stride;
  float sxy, sxy_half_off, sxy_half_off2;   // scale_x_y; 0.5*(sxy-1) (IouLoss order); (sxy-1)*0.5 (yolo_box order) -- host doubles cast to fp32
  int sxy_is_one;
  float ignore_thresh;
  int iou_aware, loss_square, match_score, has_iou_loss;
  float iou_w, aware_w;
  float aw[MAX_ANCHORS], ah[MAX_ANCHORS];
};

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float nlog(float p) { return 0.f - logf(p + 1e-9f); }
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Everything both passes need of one cell.
struct Cell {
  float sx, sy, dx, dy;          // sigmoid(x|y); decoded offsets (grid-sensitive form when scale_x_y != 1)
  float tst, tobj;               // tscale * tobj, tobj
  float x1, y1, x2, y2, pw, ph;  // predicted box (IouLoss._bbox_transform)
  float x1g, y1g, x2g, y2g;      // target box
  float iw_raw, ih_raw, iw, ih, inter, uni, iou;
  float ioup;                    // sigmoid(iou logit)
};

__device__ __forceinline__ void cell_geometry(const LossCfg& cfg, int a, int h, int w, float xl, float yl, float wl, float hl, float tx,
                                              float ty, float tw, float th, Cell& c) {
  const float S = (float)cfg.s, gx = (float)w, gy = (float)h;
  c.sx = sigm(xl); c.sy = sigm(yl);
  c.dx = cfg.sxy_is_one ? c.sx : cfg.sxy * c.sx - cfg.sxy_half_off;
  c.dy = cfg.sxy_is_one ? c.sy : cfg.sxy * c.sy - cfg.sxy_half_off;
  const float denom = (float)(cfg.s * cfg.stride);
  const float cx = (c.dx + gx) / S, cy = (c.dy + gy) / S;
  c.pw = (expf(wl) * cfg.aw[a]) / denom; c.ph = (expf(hl) * cfg.ah[a]) / denom;
  c.x1 = cx - 0.5f * c.pw; c.y1 = cy - 0.5f * c.ph; c.x2 = cx + 0.5f * c.pw; c.y2 = cy + 0.5f * c.ph;
  const float cxg = (tx + gx) / S, cyg = (ty + gy) / S;
  const float wg = (expf(tw) * cfg.aw[a]) / denom, hg = (expf(th) * cfg.ah[a]) / denom;
  c.x1g = cxg - 0.5f * wg; c.y1g = cyg - 0.5f * hg; c.x2g = cxg + 0.5f * wg; c.y2g = cyg + 0.5f * hg;
  const float x2 = fmaxf(c.x1, c.x2), y2 = fmaxf(c.y1, c.y2);
  c.iw_raw = fminf(x2, c.x2g) - fmaxf(c.x1, c.x1g);
  c.ih_raw = fminf(y2, c.y2g) - fmaxf(c.y1, c.y1g);
  c.iw = fmaxf(c.iw_raw, 0.f); c.ih = fmaxf(c.ih_raw, 0.f);
  c.inter = c.iw * c.ih;
  c.uni = (x2 - c.x1) * (y2 - c.y1) + (c.x2g - c.x1g) * (c.y2g - c.y1g) - c.inter + 1e-10f;
  c.iou = c.inter / c.uni;
}

// grid: ceil(N*A*S*chunks / WARPS) blocks of WARPS warps; warp = one 32-cell chunk of an (n, a, h) row
template <bool BWD>
__global__ void __launch_bounds__(WARPS * 32) yolo_loss_kernel(const LossCfg cfg, const float* __restrict__ out,
                                                               const float* __restrict__ tgt, const float* __restrict__ gt_box,
                                                               float* __restrict__ noobj_mask, double* __restrict__ partial,
                                                               unsigned int* __restrict__ counter, float* __restrict__ losses,
                                                               const float* __restrict__ gl, float* __restrict__ grad) {
  __shared__ double wsum[WARPS][6];
  __shared__ bool is_last;
  const int S = cfg.s, A = cfg.a, C = cfg.c;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // (a row wider than 32 cells is shared by ceil(S / 32) warps, one 32-cell chunk each: three times the warps in flight on the
  // 76 x 76 scale, whose launch is latency-bound on the per-cell ground-truth IoU loop)
  const int chunks = (S + 31) / 32;
  const int rows = cfg.n * A * S * chunks;
  const int row_c = blockIdx.x * WARPS + warp;
  const bool live = row_c < rows;
  const int chunk = live ? row_c % chunks : 0, row = row_c / chunks;
  const int n = live ? row / (A * S) : 0, a = live ? (row / S) % A : 0, h = live ? row % S : 0;
  const size_t plane = (size_t)S * S;
  const int per = 5 + C, CH = A * (per + (cfg.iou_aware ? 1 : 0));
  const float* o_n = out + (size_t)n * CH * plane;
  const float* o_a = o_n + (size_t)((cfg.iou_aware ? A : 0) + a * per) * plane + (size_t)h * S;     // channel 0 (x) of this anchor, row h
  const float* o_iou = o_n + (size_t)a * plane + (size_t)h * S;
  const float* t_a = tgt + ((size_t)(n * A + a) * (6 + C)) * plane + (size_t)h * S;
  float* g_n = BWD ? grad + (size_t)n * CH * plane : nullptr;
  float* g_a = BWD ? g_n + (size_t)((cfg.iou_aware ? A : 0) + a * per) * plane + (size_t)h * S : nullptr;
  float* g_iou = BWD ? g_n + (size_t)a * plane + (size_t)h * S : nullptr;

  // forward: the image's ground-truth boxes staged once per warp in shared memory -- the per-cell IoU loop read them from global
  // memory one dependent L2 round trip per box (~20 us of a ~55 us launch whatever the grid size)
  __shared__ float4 s_gt[BWD ? 1 : WARPS][BWD ? 1 : MAX_GT];
  if (!BWD) {
    if (live) {
      const float4* gb4 = reinterpret_cast<const float4*>(gt_box) + (size_t)n * cfg.g;
      for (int k = lane; k < cfg.g; k += 32) s_gt[BWD ? 0 : warp][BWD ? 0 : k] = __ldg(gb4 + k);
    }
    __syncwarp();
  }
  // T_row = sum_w tobj (the IoU-aware loss meets the objectness target after its own sum over w)
  float trow = 0.f;
  if (live) for (int w = lane; w < S; w += 32) trow += __ldg(t_a + 5 * plane + w);
  trow = warp_sum_f(trow);

  double acc[6] = {0., 0., 0., 0., 0., 0.};     // xy, wh, obj, cls, iou, aware
  float g_xy = 0.f, g_wh = 0.f, g_obj = 0.f, g_cls = 0.f, g_il = 0.f, g_aw = 0.f;
  if (BWD) {
    const float inv_n = 1.f / (float)cfg.n;
    g_xy = __ldg(gl + 0) * inv_n; g_wh = __ldg(gl + 1) * inv_n; g_obj = __ldg(gl + 2) * inv_n; g_cls = __ldg(gl + 3) * inv_n;
    g_il = __ldg(gl + 4) * inv_n; g_aw = __ldg(gl + 5) * inv_n;
  }
  float row_la = 0.f;                          // forward: sum_w iou * -log(ioup) of this row
  if (live) {
    for (int w = chunk * 32 + lane; w < S && w < (chunk + 1) * 32; w += 32) {
      const float xl = __ldg(o_a + w), yl = __ldg(o_a + plane + w), wl = __ldg(o_a + 2 * plane + w), hl = __ldg(o_a + 3 * plane + w);
      const float ol = __ldg(o_a + 4 * plane + w);
      const float tx = __ldg(t_a + w), ty = __ldg(t_a + plane + w), tw = __ldg(t_a + 2 * plane + w), th = __ldg(t_a + 3 * plane + w);
      const float tscale = __ldg(t_a + 4 * plane + w), tobj = __ldg(t_a + 5 * plane + w);
      Cell c;
      cell_geometry(cfg, a, h, w, xl, yl, wl, hl, tx, ty, tw, th, c);
      c.tst = tscale * tobj; c.tobj = tobj;
      c.ioup = cfg.iou_aware ? sigm(__ldg(o_iou + w)) : 0.f;
      const float p = sigm(ol);
      if (!BWD) {
        // ---- xy / wh ----
        float lx, ly;
        if (cfg.sxy_is_one) {
          lx = (tx * nlog(c.sx) + (1.f - tx) * nlog(1.f - c.sx)) * c.tst;
          ly = (ty * nlog(c.sy) + (1.f - ty) * nlog(1.f - c.sy)) * c.tst;
        } else {
          lx = fabsf(c.dx - tx) * c.tst; ly = fabsf(c.dy - ty) * c.tst;
        }
        acc[0] += (double)lx + (double)ly;
        acc[1] += (double)(fabsf(wl - tw) * c.tst) + (double)(fabsf(hl - th) * c.tst);
        // ---- IoU loss / IoU-aware loss ----
        if (cfg.has_iou_loss) {
          const float li = (cfg.loss_square ? 1.f - c.iou * c.iou : 1.f - c.iou) * cfg.iou_w;
          acc[4] += (double)(li * c.tst);
        }
        if (cfg.iou_aware) row_la += (c.iou * nlog(c.ioup)) * cfg.aware_w;
        // ---- objectness: ignore mask from the best IoU of the decoded box with any ground truth (no gradient) ----
        float mask = 0.f;
        {
          const float stride = (float)cfg.stride, Sf = (float)S;
          const float bx = (cfg.sxy * c.sx + (float)w - cfg.sxy_half_off2) * stride, by = (cfg.sxy * c.sy + (float)h - cfg.sxy_half_off2) * stride;
          const float bw = expf(wl) * cfg.aw[a], bh = expf(hl) * cfg.ah[a];
          const float px0 = (bx - bw / 2.f) / Sf / stride, py0 = (by - bh / 2.f) / Sf / stride;
          const float px1 = (bx + bw / 2.f) / Sf / stride, py1 = (by + bh / 2.f) / Sf / stride;
          const float area_a = (px1 - px0) * (py1 - py0);
          float best = -INFINITY;
          bool nan_seen = false;
          for (int k = 0; k < cfg.g; ++k) {
            const float4 b = s_gt[BWD ? 0 : warp][BWD ? 0 : k];       // cx, cy, w, h
            const float gx0 = b.x - b.z / 2.f, gy0 = b.y - b.w / 2.f, gx1 = b.x + b.z / 2.f, gy1 = b.y + b.w / 2.f;
            const float iw = fmaxf(fminf(px1, gx1) - fmaxf(px0, gx0), 0.f), ih = fmaxf(fminf(py1, gy1) - fmaxf(py0, gy0), 0.f);
            const float inter = iw * ih;
            const float iou = inter / (area_a + (gx1 - gx0) * (gy1 - gy0) - inter);
            if (iou != iou) nan_seen = true;
            best = fmaxf(best, iou);
          }
          mask = (!nan_seen && best <= cfg.ignore_thresh) ? 1.f : 0.f;       // torch.max propagates NaN; NaN <= t is false
        }
        float maxprob = 0.f;
        double lcls = 0.;
        if (tobj != 0.f || cfg.match_score) {
          // (eight classes' logits and targets are requested together: one dependent L2 round trip per class was ~30 us of
          // every launch -- the warps holding a positive cell set its duration; same summation order)
          for (int k0 = 0; k0 < C; k0 += 8) {
            float lg[8], tg[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int k = k0 + j < C ? k0 + j : C - 1;
              lg[j] = __ldg(o_a + (size_t)(5 + k) * plane + w);
              tg[j] = tobj != 0.f ? __ldg(t_a + (size_t)(6 + k) * plane + w) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (k0 + j >= C) break;
              const float pc = sigm(lg[j]);
              if (cfg.match_score) maxprob = fmaxf(maxprob, p * pc);
              if (tobj != 0.f) lcls += (double)(tg[j] * nlog(pc) + (1.f - tg[j]) * nlog(1.f - pc));
            }
          }
        }
        if (cfg.match_score && !(maxprob <= 0.25f)) mask = 0.f;
        const float noobj = (tobj > 0.f ? 0.f : 1.f) * mask;
        noobj_mask[(size_t)(n * A + a) * plane + (size_t)h * S + w] = noobj;
        acc[2] += (double)(tobj * nlog(p)) + (double)(noobj * nlog(1.f - p));
        acc[3] += lcls * (double)tobj;
      } else {
        // ================= backward: d(sum_k gl[k] * loss_k) / d(raw output) =================
        const float noobj = noobj_mask[(size_t)(n * A + a) * plane + (size_t)h * S + w];
        const float dsx = c.sx * (1.f - c.sx), dsy = c.sy * (1.f - c.sy);
        float gxl, gyl, gwl, ghl;
        if (cfg.sxy_is_one) {
          gxl = g_xy * c.tst * ((0.f - tx / (c.sx + 1e-9f)) + (1.f - tx) / (1.f - c.sx + 1e-9f)) * dsx;
          gyl = g_xy * c.tst * ((0.f - ty / (c.sy + 1e-9f)) + (1.f - ty) / (1.f - c.sy + 1e-9f)) * dsy;
        } else {
          const float ex = c.dx - tx, ey = c.dy - ty;
          gxl = g_xy * c.tst * (ex > 0.f ? 1.f : (ex < 0.f ? -1.f : 0.f)) * cfg.sxy * dsx;
          gyl = g_xy * c.tst * (ey > 0.f ? 1.f : (ey < 0.f ? -1.f : 0.f)) * cfg.sxy * dsy;
        }
        {
          const float ew = wl - tw, eh = hl - th;
          gwl = g_wh * c.tst * (ew > 0.f ? 1.f : (ew < 0.f ? -1.f : 0.f));
          ghl = g_wh * c.tst * (eh > 0.f ? 1.f : (eh < 0.f ? -1.f : 0.f));
        }
        // coefficient of d(iou): IoU loss (1 - iou^2) * w * tst, IoU-aware loss iou * -log(ioup) * w * T_row
        float k_iou = 0.f;
        if (cfg.has_iou_loss) k_iou += g_il * c.tst * cfg.iou_w * (cfg.loss_square ? -2.f * c.iou : -1.f);
        if (cfg.iou_aware) k_iou += g_aw * trow * cfg.aware_w * nlog(c.ioup);
        if (k_iou != 0.f) {
          const float bw = c.x2 - c.x1, bh = c.y2 - c.y1;          // (x2 > x1: exp() > 0)
          const bool iw_on = c.iw_raw >= 0.f, ih_on = c.ih_raw >= 0.f;
          const float dI_x1 = (iw_on && c.x1 > c.x1g) ? -c.ih : 0.f, dI_x2 = (iw_on && c.x2 < c.x2g) ? c.ih : 0.f;
          const float dI_y1 = (ih_on && c.y1 > c.y1g) ? -c.iw : 0.f, dI_y2 = (ih_on && c.y2 < c.y2g) ? c.iw : 0.f;
          const float u2 = c.uni * c.uni;
          auto diou = [&](float dI, float dP) { return (dI * c.uni - c.inter * (dP - dI)) / u2; };
          const float d_x1 = diou(dI_x1, -bh), d_x2 = diou(dI_x2, bh), d_y1 = diou(dI_y1, -bw), d_y2 = diou(dI_y2, bw);
          const float sxy_eff = cfg.sxy_is_one ? 1.f : cfg.sxy, Sf = (float)S;
          gxl += k_iou * (d_x1 + d_x2) * (sxy_eff * dsx / Sf);
          gyl += k_iou * (d_y1 + d_y2) * (sxy_eff * dsy / Sf);
          gwl += k_iou * 0.5f * (d_x2 - d_x1) * c.pw;
          ghl += k_iou * 0.5f * (d_y2 - d_y1) * c.ph;
        }
        g_a[w] = gxl; g_a[plane + w] = gyl; g_a[2 * plane + w] = gwl; g_a[3 * plane + w] = ghl;
        g_a[4 * plane + w] = g_obj * (tobj * (0.f - 1.f / (p + 1e-9f)) + noobj * (1.f / (1.f - p + 1e-9f))) * (p * (1.f - p));
        if (cfg.iou_aware)
          g_iou[w] = g_aw * trow * cfg.aware_w * c.iou * (0.f - 1.f / (c.ioup + 1e-9f)) * (c.ioup * (1.f - c.ioup));
        if (tobj != 0.f) {
          for (int k0 = 0; k0 < C; k0 += 8) {
            float lg[8], tg[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int k = k0 + j < C ? k0 + j : C - 1;
              lg[j] = __ldg(o_a + (size_t)(5 + k) * plane + w);
              tg[j] = __ldg(t_a + (size_t)(6 + k) * plane + w);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (k0 + j >= C) break;
              const float pc = sigm(lg[j]), t = tg[j];
              g_a[(size_t)(5 + k0 + j) * plane + w] = g_cls * tobj * ((0.f - t / (pc + 1e-9f)) + (1.f - t) / (1.f - pc + 1e-9f)) * (pc * (1.f - pc));
            }
          }
        } else {
          for (int k = 0; k < C; ++k) g_a[(size_t)(5 + k) * plane + w] = 0.f;
        }
      }
    }
  }
  if (BWD) return;
  // ---- forward: reduce the six sums.  IoU-aware: (sum_w la) * T_row per row ----
  row_la = warp_sum_f(row_la);
  if (lane == 0 && live && cfg.iou_aware) acc[5] += (double)(row_la * trow);
#pragma unroll
  for (int k = 0; k < 6; ++k) acc[k] = warp_sum(acc[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 6; ++k) wsum[warp][k] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double t = 0.;
    for (int w = 0; w < WARPS; ++w) t += wsum[w][threadIdx.x];
    partial[(size_t)blockIdx.x * 6 + threadIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (is_last && warp < 6) {
    // warp k folds loss k: lane-strided partial sums (independent loads in flight instead of one 700-cycle round trip per
    // block), then the shuffle tree -- a fixed order, so the result is run-to-run deterministic
    __threadfence();
    double t = 0.;
    for (unsigned b = lane; b < gridDim.x; b += 32) t += __ldcg(partial + (size_t)b * 6 + warp);
    t = warp_sum(t);
    if (lane == 0) losses[warp] += (float)(t / (double)cfg.n);     // mean over the batch of the per-image sums; scales add up in launch order
  }
  if (is_last && threadIdx.x == 0) *counter = 0u;
}

// -------------------------------------------------------------------------------------------------
// Gt2YoloTargetSingle for the whole batch.  One thread per (image, scale slot): the reference's sequential scan over the
// ground-truth boxes (later boxes overwrite earlier ones in the same cell; class bits accumulate).
// -------------------------------------------------------------------------------------------------
struct TargetCfg {
  int n, g, num_anchors, mask_len, c, gh, gw, img_h, img_w;
  float iou_thresh;
  int mask[MAX_ANCHORS];
  int anchors[2 * 16];          // all anchors (w, h) in pixels
};

__global__ void gt2yolo_target_kernel(const TargetCfg cfg, const float* __restrict__ gt_bbox, const int* __restrict__ gt_class,
                                      const float* __restrict__ gt_score, float* __restrict__ target) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= cfg.n) return;
  const size_t plane = (size_t)cfg.gh * cfg.gw;
  float* tb = target + (size_t)b * cfg.mask_len * (6 + cfg.c) * plane;
  for (int k = 0; k < cfg.g; ++k) {
    const float gx = gt_bbox[((size_t)b * cfg.g + k) * 4 + 0], gy = gt_bbox[((size_t)b * cfg.g + k) * 4 + 1];
    const float gw = gt_bbox[((size_t)b * cfg.g + k) * 4 + 2], gh = gt_bbox[((size_t)b * cfg.g + k) * 4 + 3];
    const float score = gt_score[(size_t)b * cfg.g + k];
    const int cls = gt_class[(size_t)b * cfg.g + k];
    if (gw <= 0.f || gh <= 0.f || score <= 0.f) continue;
    // best anchor over ALL anchors: IoU of (0,0,gw,gh) with (0,0,aw/W,ah/H) in double like numpy's float32 x float64 mix
    double best_iou = 0.;
    int best = -1;
    double ious[16];
    for (int an = 0; an < cfg.num_anchors; ++an) {
      const double aw = (double)cfg.anchors[2 * an] / (double)cfg.img_w, ah = (double)cfg.anchors[2 * an + 1] / (double)cfg.img_h;
      const double iw = fmin((double)gw, aw), ih = fmin((double)gh, ah);
      const double inter = iw * ih;
      const double iou = inter / ((double)gw * (double)gh + aw * ah - inter);
      ious[an] = iou;
      if (iou > best_iou) { best_iou = iou; best = an; }
    }
    const int gi = (int)(gx * (float)cfg.gw), gj = (int)(gy * (float)cfg.gh);
    if (gi < 0 || gi >= cfg.gw || gj < 0 || gj >= cfg.gh) continue;       // (the reference would raise IndexError)
    for (int slot = 0; slot < cfg.mask_len; ++slot) {
      const int an = cfg.mask[slot];
      const bool hit = an == best || (cfg.iou_thresh < 1.f && ious[an] > (double)cfg.iou_thresh);
      if (!hit) continue;
      float* t = tb + (size_t)slot * (6 + cfg.c) * plane + (size_t)gj * cfg.gw + gi;
      t[0] = gx * (float)cfg.gw - (float)gi;
      t[plane] = gy * (float)cfg.gh - (float)gj;
      t[2 * plane] = logf(gw * (float)cfg.img_w / (float)cfg.anchors[2 * an]);
      t[3 * plane] = logf(gh * (float)cfg.img_h / (float)cfg.anchors[2 * an + 1]);
      t[4 * plane] = 2.0f - gw * gh;
      t[5 * plane] = score;
      if (cls >= 0 && cls < cfg.c) t[(size_t)(6 + cls) * plane] = 1.f;
    }
  }
}

int fill_loss_cfg(LossCfg* cfg, int n, int a, int num_classes, int size, int g, const float* anchors_host, int stride, double scale_x_y,
                  float ignore_thresh, int iou_aware, int has_iou_loss, float iou_w, int loss_square, float aware_w, int match_score) {
  PPY_REQUIRE(n > 0 && a > 0 && a <= MAX_ANCHORS && num_classes > 0 && size > 0 && g >= 0 && g <= MAX_GT && anchors_host && stride > 0);
  cfg->n = n; cfg->a = a; cfg->c = num_classes; cfg->s = size; cfg->g = g; cfg->stride = stride;
  cfg->sxy = (float)scale_x_y;
  cfg->sxy_half_off = (float)(0.5 * (scale_x_y - 1.0));
  cfg->sxy_half_off2 = (float)((scale_x_y - 1.0) * 0.5);
  cfg->sxy_is_one = fabs(scale_x_y - 1.0) < 1e-10 ? 1 : 0;
  cfg->ignore_thresh = ignore_thresh;
  cfg->iou_aware = iou_aware; cfg->loss_square = loss_square; cfg->match_score = match_score; cfg->has_iou_loss = has_iou_loss;
  cfg->iou_w = iou_w; cfg->aware_w = aware_w;
  for (int i = 0; i < a; ++i) { cfg->aw[i] = anchors_host[2 * i]; cfg->ah[i] = anchors_host[2 * i + 1]; }
  return PPY_OK;
}

}  // namespace
}  // namespace ppy

extern "C" {
using namespace ppy;

int ppy_yolo_loss_workspace_bytes(int n, int a, int size) {
  return (int)(ceil_div((long long)n * a * size * ((size + 31) / 32), WARPS) * 6 * sizeof(double) + 16);
}

int ppy_yolo_loss_forward(const float* out, const float* target, const float* gt_box, int n, int a, int num_classes, int size, int g,
                          const float* anchors_host, int stride, double scale_x_y, float ignore_thresh, int iou_aware, int has_iou_loss,
                          float iou_loss_weight, int loss_square, float iou_aware_weight, int match_score, float* noobj_mask,
                          void* workspace, float* losses, ppy_stream_t s) {
  PPY_REQUIRE(out && target && noobj_mask && workspace && losses && (gt_box || g == 0));
  PPY_REQUIRE((reinterpret_cast<uintptr_t>(gt_box) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0);
  LossCfg cfg;
  int rc = fill_loss_cfg(&cfg, n, a, num_classes, size, g, anchors_host, stride, scale_x_y, ignore_thresh, iou_aware, has_iou_loss,
                         iou_loss_weight, loss_square, iou_aware_weight, match_score);
  if (rc) return rc;
  const unsigned blocks = (unsigned)ceil_div((long long)n * a * size * ((size + 31) / 32), WARPS);
  unsigned int* counter = reinterpret_cast<unsigned int*>(workspace);            // zero before the first use; the kernel re-zeroes it
  double* partial = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 16);
  yolo_loss_kernel<false><<<blocks, WARPS * 32, 0, as_stream(s)>>>(cfg, out, target, gt_box, noobj_mask, partial, counter, losses, nullptr, nullptr);
  return check_launch();
}

int ppy_yolo_loss_backward(const float* out, const float* target, int n, int a, int num_classes, int size, const float* anchors_host,
                           int stride, double scale_x_y, int iou_aware, int has_iou_loss, float iou_loss_weight, int loss_square,
                           float iou_aware_weight, const float* noobj_mask, const float* grad_losses, float* grad_out, ppy_stream_t s) {
  PPY_REQUIRE(out && target && noobj_mask && grad_losses && grad_out);
  LossCfg cfg;
  int rc = fill_loss_cfg(&cfg, n, a, num_classes, size, 0, anchors_host, stride, scale_x_y, 0.f, iou_aware, has_iou_loss, iou_loss_weight,
                         loss_square, iou_aware_weight, 0);
  if (rc) return rc;
  const unsigned blocks = (unsigned)ceil_div((long long)n * a * size * ((size + 31) / 32), WARPS);
  yolo_loss_kernel<true><<<blocks, WARPS * 32, 0, as_stream(s)>>>(cfg, out, target, nullptr, const_cast<float*>(noobj_mask), nullptr, nullptr,
                                                                   nullptr, grad_losses, grad_out);
  return check_launch();
}

int ppy_gt2yolo_target(const float* gt_bbox, const int* gt_class, const float* gt_score, int n, int g, const int* anchors_host,
                       int num_anchors, const int* mask_host, int mask_len, int num_classes, int img_h, int img_w, int downsample,
                       float iou_thresh, float* target, ppy_stream_t s) {
  PPY_REQUIRE(gt_bbox && gt_class && gt_score && anchors_host && mask_host && target);
  PPY_REQUIRE(n > 0 && g >= 0 && num_anchors > 0 && num_anchors <= 16 && mask_len > 0 && mask_len <= MAX_ANCHORS && num_classes > 0 && downsample > 0);
  TargetCfg cfg;
  cfg.n = n; cfg.g = g; cfg.num_anchors = num_anchors; cfg.mask_len = mask_len; cfg.c = num_classes;
  cfg.gh = img_h / downsample; cfg.gw = img_w / downsample; cfg.img_h = img_h; cfg.img_w = img_w; cfg.iou_thresh = iou_thresh;
  for (int i = 0; i < mask_len; ++i) { PPY_REQUIRE(mask_host[i] >= 0 && mask_host[i] < num_anchors); cfg.mask[i] = mask_host[i]; }
  for (int i = 0; i < 2 * num_anchors; ++i) cfg.anchors[i] = anchors_host[i];
  const size_t bytes = (size_t)n * mask_len * (6 + num_classes) * cfg.gh * cfg.gw * sizeof(float);
  int rc = check_cuda(cudaMemsetAsync(target, 0, bytes, as_stream(s)));
  if (rc) return rc;
  gt2yolo_target_kernel<<<(unsigned)ceil_div(n, 64), 64, 0, as_stream(s)>>>(cfg, gt_bbox, gt_class, gt_score, target);
  return check_launch();
}

}  // extern "C"
