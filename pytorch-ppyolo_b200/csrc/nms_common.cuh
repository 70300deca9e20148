// Candidate list shared by the sparse decode kernel (decode.cu) and Matrix-NMS (nms.cu).
//
// A candidate is one (box, class) score > score_threshold, stored as a 64-bit key
// (score_bits << 32 | ~flat_index), flat_index = box * num_classes + class, so that a descending key sort equals a
// stable descending score sort in the reference's row-major nonzero() order (model/matrix_nms.py:115-125).
// Every candidate is also counted in a per-image histogram of its score, binned on
// (float_bits - bits(threshold)) >> shift, from which the NMS kernel derives the bin holding the nms_top_k-th score.
//
// Workspace layout (n images): hist[n][kBins] u32 | count[n] u32, cutoff[n] i32, done[n] u32 (padded to 256 B) | keys[n][cap] u64
// (dense Matrix-NMS workspace: cap = kNmsKeyCap, followed by conf[n][num_boxes] fp32 = objectness per box)
#pragma once
#include <string.h>
#include "common.cuh"

namespace ppy {

constexpr int kBins = 4096;           // histogram bins per image
constexpr int kNmsKeyCap = 8192;      // keys per image of the dense Matrix-NMS workspace (collected candidates)

struct CandSink {
  float thr; unsigned int thr_bits; int shift;
  unsigned int* hist; unsigned int* count; unsigned long long* keys; int cap;
};

__host__ __device__ __forceinline__ int score_bin(float s, unsigned int thr_bits, int shift) {
  // s > threshold > 0 here, so the bit pattern is monotonic in s
#ifdef __CUDA_ARCH__
  unsigned int d = (__float_as_uint(s) - thr_bits) >> shift;
#else
  unsigned int b; memcpy(&b, &s, 4);
  unsigned int d = (b - thr_bits) >> shift;
#endif
  return d < (unsigned)kBins ? (int)d : kBins - 1;
}

// binning of (threshold, 1]: everything above 1 lands in the top bin
inline void score_binning(float score_threshold, unsigned int* thr_bits, int* shift) {
  float thr_pos = score_threshold > 1e-30f ? score_threshold : 1e-30f;
  unsigned int one_bits;
  float one = 1.0f;
  memcpy(thr_bits, &thr_pos, 4);
  memcpy(&one_bits, &one, 4);
  int sh = 0;
  if (one_bits > *thr_bits) while (((one_bits - *thr_bits) >> sh) >= (unsigned)kBins) ++sh;
  *shift = sh;
}

// count[n] u32 | cutoff[n] i32 (bin holding the nms_top_k-th score, written by the cutoff stage) | done[n] u32 (CTAs of the
// histogram pass that have flushed), padded to 256 B -- zeroed together with the histogram
inline size_t cand_count_bytes(int n) { return ((3 * sizeof(unsigned int) * (size_t)n + 255) / 256) * 256; }
inline size_t cand_workspace_bytes(int n, int cap) {
  return sizeof(unsigned int) * (size_t)n * kBins + cand_count_bytes(n) + sizeof(unsigned long long) * (size_t)n * cap;
}
inline CandSink cand_carve(void* ws, int n, int cap, float score_threshold) {
  CandSink c;
  char* p = reinterpret_cast<char*>(ws);
  c.hist = reinterpret_cast<unsigned int*>(p);
  p += sizeof(unsigned int) * (size_t)n * kBins;
  c.count = reinterpret_cast<unsigned int*>(p);
  p += cand_count_bytes(n);
  c.keys = reinterpret_cast<unsigned long long*>(p);
  c.cap = cap;
  c.thr = score_threshold;
  score_binning(score_threshold, &c.thr_bits, &c.shift);
  return c;
}

}  // namespace ppy
