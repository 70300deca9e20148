// Batched Matrix-NMS for sm_100a (reference model/matrix_nms.py:51-151, looped per image at
// model/head.py:462-464).  Whole batch in three launches, no host sync:
//
//   1. nms_hist_kernel    (HBM-bound scan)  per-image histogram of candidate scores (> score_threshold),
//                                           binned on (float_bits - bits(threshold)) >> shift
//   2. nms_collect_kernel (HBM-bound scan)  every CTA re-derives the cutoff bin that holds the
//                                           nms_top_k-th best score from the histogram, then appends all
//                                           candidates in bins >= cutoff as 64-bit keys
//                                           (score_bits << 32 | ~flat_index)
//   3. nms_matrix_kernel  (one CTA / image) bitonic sort of the keys in shared memory -> top nms_top_k;
//                                           boxes staged in shared memory; warp-cooperative upper-
//                                           triangular IoU + same-label test (the n x n matrix is never
//                                           materialised); column max -> compensate, column min -> decay;
//                                           post-threshold; second sort; top keep_top_k rows written.
//
// Ordering: keys sort by score descending then (box, class) flat index ascending, i.e. exactly a stable
// descending sort of the reference's row-major nonzero() order (matrix_nms.py:115-125).
// NaN semantics follow torch: min/max/clamp propagate NaN, NaN*0 = NaN, comparisons with NaN are false.
#include <math_constants.h>
#include <string.h>
#include "common.cuh"
#include "nms_common.cuh"

namespace ppy {
namespace {

constexpr int kCap = kNmsKeyCap;      // max collected candidates per image (keys in shared memory)
constexpr int kMaxN = 4000;           // max boxes entering the n x n stage (per-box arrays live in dynamic shared memory: 227 KB cap)
constexpr int kMinN = 1024;           // smallest per-box array size allocated
// dynamic shared memory of nms_matrix_kernel: kCap keys, then the per-box arrays for `nmax` boxes (41 bytes per box)
inline int matrix_smem_bytes(int nmax) { return kCap * 8 + nmax * (16 + 4 * 4 + 2 * 2 + 1) + 64; }
inline int matrix_nmax(int nms_top_k) {
  int want = nms_top_k > 0 ? nms_top_k : kMaxN;        // <= 0 ("all candidates", model/matrix_nms.py:120-125): up to kMaxN
  int n = kMinN;
  while (n < want && n < kMaxN) n <<= 1;
  return n < kMaxN ? n : kMaxN;
}
constexpr int kScanThreads = 256;
constexpr int kMatrixThreads = 1024;   // launch bound; the launch uses matrix_threads()
// CTA size of nms_matrix_kernel: its sorts are barrier-bound (ncu source view: half the kernel's samples sit at the bitonic network's
// CTA barriers) and 1024 keys are 512 compare-exchange pairs per stage -- 16 warps arrive at a barrier sooner than 32 (C5 at bs 1:
// 41.4 -> 39.4 us; 256 threads: 43.5).  Lists beyond 1024 boxes keep 1024 threads.
inline int matrix_threads(int nmax) { return nmax > 1024 ? 1024 : 512; }

struct Workspace {
  unsigned int* hist;     // [n][kBins]
  unsigned int* count;    // [n] collected keys
  int* cutoff;            // [n] bin holding the nms_top_k-th best score
  unsigned int* done;     // [n] histogram CTAs finished (the last one derives the cutoff)
  unsigned long long* keys;  // [n][kCap]
};

__device__ int find_cutoff_bin(const unsigned int* __restrict__ gh, int want, unsigned int* sh, unsigned int* part);

__global__ void __launch_bounds__(kScanThreads)
nms_hist_kernel(const float* __restrict__ scores, long long per_image, long long chunk, float thr,
                unsigned int thr_bits, int shift, unsigned int* __restrict__ hist, int want, int* __restrict__ cutoff,
                unsigned int* __restrict__ done) {
  __shared__ unsigned int sh[kBins];
  __shared__ unsigned int part[33];
  __shared__ bool last;
  for (int i = threadIdx.x; i < kBins; i += kScanThreads) sh[i] = 0;
  __syncthreads();
  const int img = blockIdx.y;
  const float* base = scores + (long long)img * per_image;
  long long lo = (long long)blockIdx.x * chunk;
  long long hi = lo + chunk < per_image ? lo + chunk : per_image;
  const bool vec = ((reinterpret_cast<uintptr_t>(base) & 15) == 0) && (chunk % 4 == 0);
  if (vec) {
    const float4* b4 = reinterpret_cast<const float4*>(base);
    long long hi4 = hi / 4;
    for (long long i0 = lo / 4 + threadIdx.x; i0 < hi4; i0 += 4 * kScanThreads) {      // four independent loads in flight
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long i = i0 + (long long)u * kScanThreads;
        v[u] = i < hi4 ? __ldg(b4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (v[u].x > thr) atomicAdd(&sh[score_bin(v[u].x, thr_bits, shift)], 1u);
        if (v[u].y > thr) atomicAdd(&sh[score_bin(v[u].y, thr_bits, shift)], 1u);
        if (v[u].z > thr) atomicAdd(&sh[score_bin(v[u].z, thr_bits, shift)], 1u);
        if (v[u].w > thr) atomicAdd(&sh[score_bin(v[u].w, thr_bits, shift)], 1u);
      }
    }
    lo = hi4 * 4;  // scalar tail (only the last chunk can have one)
  }
  for (long long i = lo + threadIdx.x; i < hi; i += kScanThreads) {
    float v = __ldg(base + i);
    if (v > thr) atomicAdd(&sh[score_bin(v, thr_bits, shift)], 1u);
  }
  __syncthreads();
  unsigned int* gh = hist + (long long)img * kBins;
  for (int i = threadIdx.x; i < kBins; i += kScanThreads)
    if (sh[i]) atomicAdd(&gh[i], sh[i]);
  // the last CTA of the image to get here derives the cutoff bin once, for every CTA of the collect pass
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(done + img, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  const int cut = find_cutoff_bin(gh, want, sh, part);
  if (threadIdx.x == 0) cutoff[img] = cut;
}

// The same for a histogram somebody else filled (the decode kernels of the whole-network path): one CTA per image.
__global__ void __launch_bounds__(kScanThreads) nms_cutoff_kernel(const unsigned int* __restrict__ hist, int want, int* __restrict__ cutoff) {
  __shared__ unsigned int sh[kBins];
  __shared__ unsigned int part[33];
  const int cut = find_cutoff_bin(hist + (long long)blockIdx.x * kBins, want, sh, part);
  if (threadIdx.x == 0) cutoff[blockIdx.x] = cut;
}

// Cutoff bin: the largest bin b such that count(bins >= b) >= want; 0 if fewer candidates than `want`.
__device__ int find_cutoff_bin(const unsigned int* __restrict__ gh, int want, unsigned int* sh /*kBins*/,
                               unsigned int* part /*blockDim/32 + 1*/) {
  const int T = blockDim.x;
  const int per = kBins / T;  // bins per thread, contiguous, thread 0 owns the TOP bins
  unsigned int local = 0;
  for (int k = 0; k < per; ++k) {
    int bin = kBins - 1 - (threadIdx.x * per + k);
    unsigned int v = __ldcg(gh + bin);          // L2 read: other CTAs of this launch may have just added to it
    sh[bin] = v;
    local += v;
  }
  // inclusive scan of `local` over threads (top bins first)
  unsigned int incl = local;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) part[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    unsigned int v = lane < (T >> 5) ? part[lane] : 0;
    for (int o = 1; o < 32; o <<= 1) {
      unsigned int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    part[lane] = v;  // inclusive warp totals
  }
  __syncthreads();
  unsigned int before = incl - local + (wid ? part[wid - 1] : 0);  // candidates in bins above this thread's
  __shared__ int s_cut;
  if (threadIdx.x == 0) s_cut = 0;
  __syncthreads();
  if (before < (unsigned)want && before + local >= (unsigned)want) {
    unsigned int acc = before;
    for (int k = 0; k < per; ++k) {
      int bin = kBins - 1 - (threadIdx.x * per + k);
      acc += sh[bin];
      if (acc >= (unsigned)want) { s_cut = bin; break; }
    }
  }
  __syncthreads();
  return s_cut;
}

__global__ void __launch_bounds__(kScanThreads)
nms_collect_kernel(const float* __restrict__ scores, long long per_image, long long chunk, float thr,
                   unsigned int thr_bits, int shift, const int* __restrict__ cutoff, unsigned int* __restrict__ count,
                   unsigned long long* __restrict__ keys) {
  const int img = blockIdx.y;
  const int cut = cutoff[img];
  const float* base = scores + (long long)img * per_image;
  unsigned long long* out = keys + (long long)img * kCap;
  unsigned int* cnt = count + img;
  long long lo = (long long)blockIdx.x * chunk;
  long long hi = lo + chunk < per_image ? lo + chunk : per_image;
  auto emit = [&](float v, long long idx) {
    if (v > thr && score_bin(v, thr_bits, shift) >= cut) {
      unsigned int slot = atomicAdd(cnt, 1u);
      if (slot < (unsigned)kCap)
        out[slot] = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)idx);
    }
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(base) & 15) == 0) && (chunk % 4 == 0);
  if (vec) {
    const float4* b4 = reinterpret_cast<const float4*>(base);
    long long hi4 = hi / 4;
    for (long long i0 = lo / 4 + threadIdx.x; i0 < hi4; i0 += 4 * kScanThreads) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long i = i0 + (long long)u * kScanThreads;
        v[u] = i < hi4 ? __ldg(b4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long i = i0 + (long long)u * kScanThreads;
        emit(v[u].x, 4 * i); emit(v[u].y, 4 * i + 1); emit(v[u].z, 4 * i + 2); emit(v[u].w, 4 * i + 3);
      }
    }
    lo = hi4 * 4;
  }
  for (long long i = lo + threadIdx.x; i < hi; i += kScanThreads) emit(__ldg(base + i), i);
}

// Collect with box pruning (whole-network path): conf[img][box] is the objectness the scores were multiplied by, so a box
// whose conf falls below the cutoff bin cannot hold a candidate.  One thread per (box, 4 consecutive classes): the conf
// test is an L1-resident load shared by the box's threads, and only surviving boxes have their scores read (one 16-byte
// load per thread).  Same keys as nms_collect_kernel.  Needs num_classes % 4 == 0.
__global__ void __launch_bounds__(kScanThreads)
nms_collect_pruned_kernel(const float* __restrict__ scores, const float* __restrict__ conf, int num_boxes, int num_classes,
                          float thr, unsigned int thr_bits, int shift, const int* __restrict__ cutoff,
                          unsigned int* __restrict__ count, unsigned long long* __restrict__ keys) {
  const int img = blockIdx.y;
  const int cut = cutoff[img];
  const float4* sc = reinterpret_cast<const float4*>(scores + (long long)img * num_boxes * num_classes);
  const float* cf = conf + (long long)img * num_boxes;
  unsigned long long* out = keys + (long long)img * kCap;
  unsigned int* cnt = count + img;
  const int qpb = num_classes / 4;
  const int quads = num_boxes * qpb;
  auto emit = [&](float v, unsigned int idx) {
    if (v > thr && score_bin(v, thr_bits, shift) >= cut) {
      const unsigned int slot = atomicAdd(cnt, 1u);
      if (slot < (unsigned)kCap) out[slot] = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
    }
  };
  for (int q = blockIdx.x * kScanThreads + threadIdx.x; q < quads; q += gridDim.x * kScanThreads) {
    const int b = q / qpb;
    const float c = __ldg(cf + b);
    if (!(c > thr) || score_bin(c, thr_bits, shift) < cut) continue;
    const float4 v = __ldg(sc + q);
    const unsigned int idx = 4u * (unsigned int)q;          // = b * num_classes + 4 * (q % qpb)
    emit(v.x, idx); emit(v.y, idx + 1); emit(v.z, idx + 2); emit(v.w, idx + 3);
  }
}

// descending bitonic sort of `len` (power of two) 64-bit keys in shared memory.
// The network is latency-bound (145 dependent stages for the kernel's three sorts), so every stage whose partner distance is
// <= 32 runs in REGISTERS: a warp owns 64 consecutive keys, lane l holds keys l and l + 32; distance 32 is a compare-exchange
// inside the thread, distances 16..1 are shuffles.  Merge sizes 2..64 never touch shared memory; for larger sizes only the
// distances >= 64 are shared-memory stages with a CTA barrier (10 for 1024 keys), followed by one register phase.
__device__ __forceinline__ unsigned long long key_pick(unsigned long long x, unsigned long long y, bool want_max) {
  return want_max ? (x > y ? x : y) : (x < y ? x : y);
}
// distances 16..1 (or from `first` down) of merge size `size` on this lane's two keys (indices i0, i0 + 32)
__device__ __forceinline__ void bitonic_shuffle_steps(unsigned long long& a, unsigned long long& b, int i0, int size, int first) {
  const int lane = threadIdx.x & 31;
  const bool desc = (i0 & size) == 0;       // (i0 and i0 + 32 share that bit for every size >= 64; for smaller sizes see the caller)
  for (int stride = first; stride > 0; stride >>= 1) {
    const unsigned long long pa = __shfl_xor_sync(0xffffffffu, a, stride), pb = __shfl_xor_sync(0xffffffffu, b, stride);
    const bool lo = (lane & stride) == 0;    // this lane holds the lower index of the pair
    a = key_pick(a, pa, lo == desc);
    b = key_pick(b, pb, lo == desc);
  }
}
__device__ void bitonic_sort_desc(unsigned long long* k, int len) {
  __syncthreads();
  if (len < 64) {                            // tiny lists: plain shared-memory network
    for (int size = 2; size <= len; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int t = threadIdx.x; t < (len >> 1); t += blockDim.x) {
          int lo = 2 * t - (t & (stride - 1));
          int hi = lo + stride;
          bool desc = ((lo & size) == 0);
          unsigned long long a = k[lo], b = k[hi];
          if ((a < b) == desc) { k[lo] = b; k[hi] = a; }
        }
        __syncthreads();
      }
    }
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // merge sizes 2 .. 64, all in registers
  for (int base = warp * 64; base < len; base += nw * 64) {
    const int i0 = base + lane;
    unsigned long long a = k[i0], b = k[i0 + 32];
    for (int size = 2; size <= 32; size <<= 1) {
      // (sizes < 64: keys l and l + 32 differ in bit 5 only, so (i & size) is the same for both when size <= 16; for size == 32
      // it differs: handled by flipping the direction of b)
      const bool da = (i0 & size) == 0, db = ((i0 + 32) & size) == 0;
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        const unsigned long long pa = __shfl_xor_sync(0xffffffffu, a, stride), pb = __shfl_xor_sync(0xffffffffu, b, stride);
        const bool lo = (lane & stride) == 0;
        a = key_pick(a, pa, lo == da);
        b = key_pick(b, pb, lo == db);
      }
    }
    {                                        // size 64: distance 32 inside the thread, then 16..1
      const bool desc = (i0 & 64) == 0;
      const unsigned long long mx = a > b ? a : b, mn = a > b ? b : a;
      a = desc ? mx : mn; b = desc ? mn : mx;
      bitonic_shuffle_steps(a, b, i0, 64, 16);
    }
    k[i0] = a; k[i0 + 32] = b;
  }
  for (int size = 128; size <= len; size <<= 1) {
    for (int stride = size >> 1; stride >= 64; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (len >> 1); t += blockDim.x) {
        int lo = 2 * t - (t & (stride - 1));
        int hi = lo + stride;
        bool desc = ((lo & size) == 0);
        unsigned long long a = k[lo], b = k[hi];
        if ((a < b) == desc) { k[lo] = b; k[hi] = a; }
      }
    }
    __syncthreads();
    for (int base = warp * 64; base < len; base += nw * 64) {
      const int i0 = base + lane;
      unsigned long long a = k[i0], b = k[i0 + 32];
      const bool desc = (i0 & size) == 0;
      const unsigned long long mx = a > b ? a : b, mn = a > b ? b : a;
      a = desc ? mx : mn; b = desc ? mn : mx;
      bitonic_shuffle_steps(a, b, i0, size, 16);
      k[i0] = a; k[i0 + 32] = b;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float nan_min(float a, float b) { return (a != a || b != b) ? CUDART_NAN_F : fminf(a, b); }
__device__ __forceinline__ float nan_max(float a, float b) { return (a != a || b != b) ? CUDART_NAN_F : fmaxf(a, b); }
__device__ __forceinline__ float clamp_min0(float v) { return v < 0.f ? 0.f : v; }  // torch.clamp(min=0): NaN stays NaN

// jaccard of two xyxy boxes with exactly the reference's fp32 operation order (matrix_nms.py:15-47)
__device__ __forceinline__ float box_iou(const float4 a, const float4 b) {
  float iw = clamp_min0(__fsub_rn(nan_min(a.z, b.z), nan_max(a.x, b.x)));
  float ih = clamp_min0(__fsub_rn(nan_min(a.w, b.w), nan_max(a.y, b.y)));
  float inter = __fmul_rn(iw, ih);
  float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  float area_b = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
  return __fdiv_rn(inter, uni);
}

// One CTA per image.  `hist` == nullptr: `keys` already holds only the candidates at/above the cutoff bin
// (nms_collect_kernel).  Otherwise `keys` holds EVERY candidate (sparse decode, decode.cu) and the CTA first derives the
// cutoff bin from the histogram and filters the list into shared memory.
__global__ void __launch_bounds__(kMatrixThreads)
nms_matrix_kernel(const float* __restrict__ boxes, int num_boxes, int num_classes, const unsigned int* __restrict__ count,
                  const unsigned long long* __restrict__ keys, int key_stride, const unsigned int* __restrict__ hist,
                  unsigned int thr_bits, int shift, int nms_top_k, int keep_top_k, float post_thr,
                  int use_gaussian, float sigma, float* __restrict__ out, int* __restrict__ counts, int nmax) {
  extern __shared__ unsigned long long s_keys[];           // kCap keys, later reused for the other sorts
  // per-box arrays for up to nmax boxes, carved from the dynamic allocation behind the keys
  float4* s_box = reinterpret_cast<float4*>(s_keys + kCap);
  float* s_score = reinterpret_cast<float*>(s_box + nmax);
  int* s_label = reinterpret_cast<int*>(s_score + nmax);
  float* s_comp = reinterpret_cast<float*>(s_label + nmax);
  float* s_new = s_comp + nmax;
  unsigned short* s_order = reinterpret_cast<unsigned short*>(s_new + nmax);    // box indices grouped by label (ascending index inside a label)
  unsigned short* s_gstart = s_order + nmax;                                      // first position of the label group a position belongs to
  unsigned char* s_odd = reinterpret_cast<unsigned char*>(s_gstart + nmax);       // box whose IoU can be NaN/inf (area not positive finite)
  __shared__ unsigned int s_part[33];
  __shared__ int s_nan_from;   // largest i with NaN compensate (poisons every column j <= i), -1 if none
  __shared__ int s_kept;
  __shared__ int s_any_odd;
  __shared__ int s_maxgrp;     // longest same-label prefix a column has to scan
  __shared__ unsigned int s_m;
  const int img = blockIdx.x;
  const int nthreads = blockDim.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = nthreads >> 5;
  const unsigned int total = count[img];
  float* my_out = out + (long long)img * keep_top_k * 6;
  if (total > (unsigned)key_stride) {   // candidate list overflowed -- flagged, see DESIGN.md
    if (tid == 0) counts[img] = -2;
    return;
  }
  if (total == 0) { if (tid == 0) counts[img] = 0; return; }
  const unsigned long long* gk = keys + (long long)img * key_stride;
  int m;
  if (hist != nullptr) {
    const int want = nms_top_k > 0 ? nms_top_k : kCap + 1;
    const int cut = find_cutoff_bin(hist + (long long)img * kBins, want, reinterpret_cast<unsigned int*>(s_keys), s_part);
    if (tid == 0) s_m = 0;
    __syncthreads();
    for (unsigned int i = tid; i < total; i += nthreads) {
      const unsigned long long k = gk[i];
      if (score_bin(__uint_as_float((unsigned int)(k >> 32)), thr_bits, shift) >= cut) {
        const unsigned int slot = atomicAdd(&s_m, 1u);
        if (slot < (unsigned)kCap) s_keys[slot] = k;
      }
    }
    __syncthreads();
    if (s_m > (unsigned)kCap) {         // cutoff bin too crowded (mass ties)
      if (tid == 0) counts[img] = -2;
      return;
    }
    m = (int)s_m;
  } else {
    m = (int)total;
    for (int i = tid; i < m; i += nthreads) s_keys[i] = gk[i];
  }
  int len = 1; while (len < m) len <<= 1;
  for (int i = m + tid; i < len; i += nthreads) s_keys[i] = 0ull;
  bitonic_sort_desc(s_keys, len);
  int n = m;
  if (nms_top_k > 0 && n > nms_top_k) n = nms_top_k;
  if (n > nmax) { if (tid == 0) counts[img] = -3; return; }
  if (tid == 0) { s_nan_from = -1; s_kept = 0; s_any_odd = 0; s_maxgrp = 0; }
  __syncthreads();
  const float4* gb = reinterpret_cast<const float4*>(boxes) + (long long)img * num_boxes;
  for (int i = tid; i < n; i += nthreads) {
    unsigned long long k = s_keys[i];
    unsigned int flat = 0xFFFFFFFFu - (unsigned int)(k & 0xFFFFFFFFull);
    s_score[i] = __uint_as_float((unsigned int)(k >> 32));
    s_label[i] = (int)(flat % (unsigned)num_classes);
    const float4 bx = __ldg(gb + flat / (unsigned)num_classes);
    s_box[i] = bx;
    const float area = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
    const bool odd = !(area > 0.f && area < CUDART_INF_F);
    s_odd[i] = odd;
    if (odd) s_any_odd = 1;
  }
  __syncthreads();
  const float neg_sigma = -1.f * sigma;
  if (!s_any_odd) {
    // ---- every area is positive and finite: IoU is never NaN and pairs with different labels (or disjoint boxes)
    // contribute exactly +0 to compensate and a factor >= 1 to the decay min, so only same-label pairs matter.
    // Group the boxes by label (sort of (label, index)) and let each warp walk just its column's group prefix.
    int len2 = 1; while (len2 < n) len2 <<= 1;
    for (int i = tid; i < len2; i += nthreads)
      s_keys[i] = i < n ? (((unsigned long long)(0xFFFFFFFFu - (unsigned int)s_label[i]) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)i)) : 0ull;
    bitonic_sort_desc(s_keys, len2);
    for (int p = tid; p < n; p += nthreads) s_order[p] = (unsigned short)(0xFFFFFFFFu - (unsigned int)(s_keys[p] & 0xFFFFFFFFull));
    __syncthreads();
    for (int p = tid; p < n; p += nthreads) {
      // first position of p's label group: s_order is sorted by (label, index), so a binary search over [0, p] finds it in
      // ~log2(n) dependent shared-memory reads (walking back one position at a time cost up to the group length: 13 % of the
      // kernel when a few labels hold most of the 500 boxes)
      const int lp = s_label[s_order[p]];
      int lo = 0, hi = p;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_label[s_order[mid]] < lp) lo = mid + 1; else hi = mid;
      }
      const int q = lo;
      s_gstart[p] = (unsigned short)q;
      if (p - q > 32) atomicMax(&s_maxgrp, p - q);
    }
    __syncthreads();
    if (s_maxgrp == 0) {
      // every label group is short (<= 33 boxes): one THREAD per column walks its prefix -- no shuffle reductions, all columns
      // at once (the warp-per-column form below costs ~450 dependent cycles per column and 16 columns per warp)
      for (int p = tid; p < n; p += nthreads) {
        const int j = s_order[p];
        const float4 bj = s_box[j];
        float mx = 0.f;
        for (int q = s_gstart[p]; q < p; ++q) {
          const float4 bi = s_box[s_order[q]];
          const float iw = __fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x)), ih = __fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y));
          if (iw <= 0.f || ih <= 0.f) continue;
          mx = fmaxf(mx, box_iou(bi, bj));
        }
        s_comp[j] = mx;
      }
      __syncthreads();
      for (int p = tid; p < n; p += nthreads) {
        const int j = s_order[p];
        const float4 bj = s_box[j];
        float mn = j > 0 ? 1.f : CUDART_INF_F;
        for (int q = s_gstart[p]; q < p; ++q) {
          const int i = s_order[q];
          const float4 bi = s_box[i];
          const float iw = __fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x)), ih = __fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y));
          if (iw <= 0.f || ih <= 0.f) continue;
          const float d = box_iou(bi, bj);
          const float c = s_comp[i];
          float e;
          if (use_gaussian) e = __fdiv_rn(expf(__fmul_rn(neg_sigma, __fmul_rn(d, d))), expf(__fmul_rn(neg_sigma, __fmul_rn(c, c))));
          else e = __fdiv_rn(__fsub_rn(1.f, d), __fsub_rn(1.f, c));
          mn = nan_min(mn, e);
        }
        const float cj = s_comp[j];
        const float self = use_gaussian ? __fdiv_rn(1.f, expf(__fmul_rn(neg_sigma, __fmul_rn(cj, cj)))) : __fdiv_rn(1.f, __fsub_rn(1.f, cj));
        mn = nan_min(mn, self);
        s_new[j] = __fmul_rn(s_score[j], mn);
      }
    } else {
    for (int p = wid; p < n; p += nwarps) {
      const int j = s_order[p];
      const float4 bj = s_box[j];
      float mx = 0.f;
      for (int q = s_gstart[p] + lane; q < p; q += 32) {
        const float4 bi = s_box[s_order[q]];
        const float iw = __fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x)), ih = __fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y));
        if (iw <= 0.f || ih <= 0.f) continue;
        mx = fmaxf(mx, box_iou(bi, bj));
      }
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (lane == 0) s_comp[j] = mx;
    }
    __syncthreads();
    for (int p = wid; p < n; p += nwarps) {
      const int j = s_order[p];
      const float4 bj = s_box[j];
      float mn = j > 0 ? 1.f : CUDART_INF_F;
      for (int q = s_gstart[p] + lane; q < p; q += 32) {
        const int i = s_order[q];
        const float4 bi = s_box[i];
        const float iw = __fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x)), ih = __fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y));
        if (iw <= 0.f || ih <= 0.f) continue;
        const float d = box_iou(bi, bj);
        const float c = s_comp[i];
        float e;
        if (use_gaussian) e = __fdiv_rn(expf(__fmul_rn(neg_sigma, __fmul_rn(d, d))), expf(__fmul_rn(neg_sigma, __fmul_rn(c, c))));
        else e = __fdiv_rn(__fsub_rn(1.f, d), __fsub_rn(1.f, c));
        mn = nan_min(mn, e);
      }
      for (int o = 16; o > 0; o >>= 1) mn = nan_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      if (lane == 0) {
        const float cj = s_comp[j];
        const float self = use_gaussian ? __fdiv_rn(1.f, expf(__fmul_rn(neg_sigma, __fmul_rn(cj, cj)))) : __fdiv_rn(1.f, __fsub_rn(1.f, cj));
        mn = nan_min(mn, self);
        s_new[j] = __fmul_rn(s_score[j], mn);
      }
    }
    }
  } else {
  // compensate[j] = max_i (iou*same)[i][j] over the strict upper triangle (matrix_nms.py:67-78); warp per column
  for (int j = wid; j < n; j += nwarps) {
    const float4 bj = s_box[j];
    const int lj = s_label[j];
    const bool oddj = s_odd[j];
    float mx = 0.f;
    for (int i = lane; i < j; i += 32) {
      const bool same = s_label[i] == lj;
      if (!oddj && !s_odd[i]) {
        // both areas positive and finite: no NaN possible; different labels or disjoint boxes give exactly +0
        if (!same) continue;
        const float4 bi = s_box[i];
        const float iw = __fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x)), ih = __fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y));
        if (iw <= 0.f || ih <= 0.f) continue;
        mx = fmaxf(mx, box_iou(bi, bj));
      } else {
        mx = nan_max(mx, __fmul_rn(box_iou(s_box[i], bj), same ? 1.f : 0.f));
      }
    }
    for (int o = 16; o > 0; o >>= 1) mx = nan_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) {
      s_comp[j] = mx;
      if (mx != mx) atomicMax(&s_nan_from, j);
    }
  }
  __syncthreads();
  // decay[j] = min_i f(d[i][j]) / f(compensate[i])  (:85-93).  Rows i >= j have d = 0 and contribute
  // 1/f(comp_i) >= 1 >= row 0's term, so only their NaNs matter (s_nan_from); row i < j evaluated exactly.
  const int nan_from = s_nan_from;
  for (int j = wid; j < n; j += nwarps) {
    const float4 bj = s_box[j];
    const int lj = s_label[j];
    const bool oddj = s_odd[j];
    // row 0 (comp_0 == 0) contributes (1 - d_0j) <= 1, exactly 1 when skipped below: start the min at 1 for j > 0.
    // Rows with d == 0 contribute 1/f(comp_i) >= 1 and can be skipped; a NaN compensate anywhere makes EVERY
    // column NaN (it sits in a full row of the decay matrix), which s_nan_from handles globally below.
    float mn = j > 0 ? 1.f : CUDART_INF_F;
    for (int i = lane; i < j; i += 32) {
      const bool same = s_label[i] == lj;
      float d;
      if (!oddj && !s_odd[i]) {
        if (!same) continue;
        const float4 bi = s_box[i];
        const float iw = __fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x)), ih = __fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y));
        if (iw <= 0.f || ih <= 0.f) continue;
        d = box_iou(bi, bj);
      } else {
        d = __fmul_rn(box_iou(s_box[i], bj), same ? 1.f : 0.f);
      }
      const float c = s_comp[i];
      float e;
      if (use_gaussian) e = __fdiv_rn(expf(__fmul_rn(neg_sigma, __fmul_rn(d, d))), expf(__fmul_rn(neg_sigma, __fmul_rn(c, c))));
      else e = __fdiv_rn(__fsub_rn(1.f, d), __fsub_rn(1.f, c));
      mn = nan_min(mn, e);
    }
    for (int o = 16; o > 0; o >>= 1) mn = nan_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    if (lane == 0) {
      // rows i >= j: term (f(0)/f(comp_i)); row j itself always exists
      float cj = s_comp[j];
      float self = use_gaussian ? __fdiv_rn(1.f, expf(__fmul_rn(neg_sigma, __fmul_rn(cj, cj)))) : __fdiv_rn(1.f, __fsub_rn(1.f, cj));
      mn = nan_min(mn, self);   // j == 0: comp[0] == 0 -> exactly 1
      if (nan_from >= 0) mn = CUDART_NAN_F;
      s_new[j] = __fmul_rn(s_score[j], mn);
    }
  }
  }
  __syncthreads();
  // post threshold (>=, NaN fails; :132) then stable descending sort by decayed score (:140-145)
  int len3 = 1; while (len3 < n) len3 <<= 1;
  for (int i = tid; i < len3; i += nthreads) {
    unsigned long long k = 0ull;
    if (i < n) {
      float v = s_new[i];
      if (v >= post_thr) {
        // sign-magnitude -> monotonic unsigned (scores are >= 0 in practice; this keeps negatives ordered too)
        unsigned int b = __float_as_uint(v);
        b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
        k = ((unsigned long long)b << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)i);
        atomicAdd(&s_kept, 1);
      }
    }
    s_keys[i] = k;
  }
  bitonic_sort_desc(s_keys, len3);
  int kept = s_kept;
  if (kept > keep_top_k) kept = keep_top_k;
  for (int r = tid; r < kept; r += nthreads) {
    int i = (int)(0xFFFFFFFFu - (unsigned int)(s_keys[r] & 0xFFFFFFFFull));
    float4 b = s_box[i];
    float* row = my_out + r * 6;
    row[0] = (float)s_label[i]; row[1] = s_new[i]; row[2] = b.x; row[3] = b.y; row[4] = b.z; row[5] = b.w;
  }
  if (tid == 0) counts[img] = kept;
}

__global__ void pairwise_iou_kernel(const float4* __restrict__ a, int na, const float4* __restrict__ b, int nb,
                                    float* __restrict__ out) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y;
  if (j < nb && i < na) out[(long long)i * nb + j] = box_iou(__ldg(a + i), __ldg(b + j));
}

Workspace carve(void* ws, int n) {
  Workspace w;
  char* p = reinterpret_cast<char*>(ws);
  w.hist = reinterpret_cast<unsigned int*>(p);
  p += sizeof(unsigned int) * (size_t)n * kBins;
  w.count = reinterpret_cast<unsigned int*>(p);
  w.cutoff = reinterpret_cast<int*>(w.count + n);
  w.done = w.count + 2 * n;
  p += cand_count_bytes(n);
  w.keys = reinterpret_cast<unsigned long long*>(p);
  return w;
}

size_t workspace_bytes(int n, int num_boxes) {
  return sizeof(unsigned int) * (size_t)n * kBins + cand_count_bytes(n) +
         sizeof(unsigned long long) * (size_t)n * kCap + sizeof(float) * (size_t)n * num_boxes;
}

}  // namespace
}  // namespace ppy

extern "C" {

int ppy_matrix_nms_workspace_bytes(int n, int num_boxes, int num_classes, size_t* bytes) {
  PPY_REQUIRE(bytes && n > 0 && num_boxes > 0 && num_classes > 0);
  *bytes = ppy::workspace_bytes(n, num_boxes);
  return PPY_OK;
}

static int matrix_nms_dense(const float* boxes, const float* scores, int n, int num_boxes, int num_classes,
                            float score_threshold, float post_threshold, int nms_top_k, int keep_top_k,
                            int use_gaussian, float gaussian_sigma, float* out, int* counts, void* workspace,
                            size_t workspace_bytes, ppy_stream_t s, bool have_hist) {
  using namespace ppy;
  PPY_REQUIRE(boxes && scores && out && counts && workspace);
  PPY_REQUIRE(n > 0 && num_boxes > 0 && num_classes > 0 && keep_top_k > 0);
  PPY_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15) == 0);
  PPY_REQUIRE(nms_top_k <= kMaxN);
  PPY_REQUIRE((long long)num_boxes * num_classes < 0xFFFFFFFFll);
  if (workspace_bytes < ppy::workspace_bytes(n, num_boxes)) return PPY_ERR_WORKSPACE;
  cudaStream_t st = as_stream(s);
  Workspace w = carve(workspace, n);
  // histogram + count live at the head of the workspace, contiguous
  int rc = PPY_OK;
  if (!have_hist) {
    rc = check_cuda(cudaMemsetAsync(w.hist, 0, reinterpret_cast<char*>(w.keys) - reinterpret_cast<char*>(w.hist), st));
    if (rc) return rc;
  }
  unsigned int thr_bits;
  int shift;
  score_binning(score_threshold, &thr_bits, &shift);
  const long long per_image = (long long)num_boxes * num_classes;
  // enough CTAs to saturate HBM: ~4 per SM over the whole batch, each CTA >= 4K scores (four 16-byte loads per thread), chunk % 4 == 0
  long long gx = ceil_div(148 * 4, n);
  long long max_gx = ceil_div(per_image, 4096);
  if (gx > max_gx) gx = max_gx;
  if (gx < 1) gx = 1;
  long long chunk = ceil_div(ceil_div(per_image, gx), 4) * 4;
  gx = ceil_div(per_image, chunk);
  dim3 grid((unsigned)gx, (unsigned)n);
  int want = nms_top_k > 0 ? nms_top_k : kCap + 1;   // <=0: take everything (cutoff bin 0)
  if (!have_hist) {
    nms_hist_kernel<<<grid, kScanThreads, 0, st>>>(scores, per_image, chunk, score_threshold, thr_bits, shift, w.hist, want, w.cutoff, w.done);
    if ((rc = check_launch())) return rc;
  } else {
    nms_cutoff_kernel<<<n, kScanThreads, 0, st>>>(w.hist, want, w.cutoff);
    if ((rc = check_launch())) return rc;
  }
  if (have_hist && num_classes % 4 == 0 && (reinterpret_cast<uintptr_t>(scores) & 15) == 0) {
    // decode left the objectness per box behind the keys: prune whole boxes (score <= conf)
    const float* conf = reinterpret_cast<const float*>(w.keys + (size_t)n * kCap);
    const long long quads = (long long)num_boxes * (num_classes / 4);
    long long cx = ceil_div(148 * 6, n);
    if (cx > ceil_div(quads, kScanThreads)) cx = ceil_div(quads, kScanThreads);
    if (cx < 1) cx = 1;
    dim3 cgrid((unsigned)cx, (unsigned)n);
    nms_collect_pruned_kernel<<<cgrid, kScanThreads, 0, st>>>(scores, conf, num_boxes, num_classes, score_threshold, thr_bits, shift,
                                                              w.cutoff, w.count, w.keys);
  } else {
    nms_collect_kernel<<<grid, kScanThreads, 0, st>>>(scores, per_image, chunk, score_threshold, thr_bits, shift, w.cutoff,
                                                      w.count, w.keys);
  }
  if ((rc = check_launch())) return rc;
  static DeviceOnce attr_once;
  const int nmax = matrix_nmax(nms_top_k);
  const int smem = matrix_smem_bytes(nmax);
  if (attr_once.first()) {
    rc = check_cuda(cudaFuncSetAttribute(nms_matrix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, matrix_smem_bytes(kMaxN)));
    if (rc) return rc;
  }
  nms_matrix_kernel<<<n, matrix_threads(nmax), smem, st>>>(boxes, num_boxes, num_classes, w.count, w.keys, kCap, nullptr, 0u, 0,
                                                     nms_top_k, keep_top_k, post_threshold, use_gaussian, gaussian_sigma,
                                                     out, counts, nmax);
  return check_launch();
}

int ppy_matrix_nms_batched(const float* boxes, const float* scores, int n, int num_boxes, int num_classes,
                           float score_threshold, float post_threshold, int nms_top_k, int keep_top_k,
                           int use_gaussian, float gaussian_sigma, float* out, int* counts, void* workspace,
                           size_t workspace_bytes, ppy_stream_t s) {
  return matrix_nms_dense(boxes, scores, n, num_boxes, num_classes, score_threshold, post_threshold, nms_top_k, keep_top_k,
                          use_gaussian, gaussian_sigma, out, counts, workspace, workspace_bytes, s, false);
}

int ppy_matrix_nms_batched_hist(const float* boxes, const float* scores, int n, int num_boxes, int num_classes,
                                float score_threshold, float post_threshold, int nms_top_k, int keep_top_k,
                                int use_gaussian, float gaussian_sigma, float* out, int* counts, void* workspace,
                                size_t workspace_bytes, ppy_stream_t s) {
  return matrix_nms_dense(boxes, scores, n, num_boxes, num_classes, score_threshold, post_threshold, nms_top_k, keep_top_k,
                          use_gaussian, gaussian_sigma, out, counts, workspace, workspace_bytes, s, true);
}

int ppy_matrix_nms_candidates(const float* boxes, int n, int num_boxes, int num_classes, float score_threshold,
                              float post_threshold, int nms_top_k, int keep_top_k, int use_gaussian, float gaussian_sigma,
                              float* out, int* counts, void* workspace, int cap, ppy_stream_t s) {
  using namespace ppy;
  PPY_REQUIRE(boxes && out && counts && workspace);
  PPY_REQUIRE(n > 0 && num_boxes > 0 && num_classes > 0 && keep_top_k > 0 && cap > 0 && score_threshold > 0.f);
  PPY_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15) == 0);
  PPY_REQUIRE(nms_top_k <= kMaxN);
  PPY_REQUIRE((long long)num_boxes * num_classes < 0xFFFFFFFFll);
  const CandSink c = cand_carve(workspace, n, cap, score_threshold);
  static DeviceOnce attr_once;
  const int nmax = matrix_nmax(nms_top_k);
  const int smem = matrix_smem_bytes(nmax);
  if (attr_once.first()) {
    int rc = check_cuda(cudaFuncSetAttribute(nms_matrix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, matrix_smem_bytes(kMaxN)));
    if (rc) return rc;
  }
  nms_matrix_kernel<<<n, matrix_threads(nmax), smem, as_stream(s)>>>(boxes, num_boxes, num_classes, c.count, c.keys, cap, c.hist,
                                                               c.thr_bits, c.shift, nms_top_k, keep_top_k, post_threshold,
                                                               use_gaussian, gaussian_sigma, out, counts, nmax);
  return check_launch();
}

int ppy_pairwise_iou(const float* a, int na, const float* b, int nb, float* out, ppy_stream_t s) {
  using namespace ppy;
  PPY_REQUIRE(a && b && out && na > 0 && nb > 0);
  PPY_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0);
  dim3 grid((unsigned)ceil_div(nb, 128), (unsigned)na);
  pairwise_iou_kernel<<<grid, 128, 0, as_stream(s)>>>(reinterpret_cast<const float4*>(a), na,
                                                      reinterpret_cast<const float4*>(b), nb, out);
  return check_launch();
}

}  // extern "C"
