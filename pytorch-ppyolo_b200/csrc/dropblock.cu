// DropBlock (reference model/custom_layers.py:293-342), the training-time block dropout of the YOLOv3 head, as three small
// HBM-bound kernels with a counter-based RNG (no host round trip, CUDA-graph capturable):
//   seeds  Bernoulli(gamma) draws, one Philox4x32-10 call per four consecutive logical (n,c,h,w) elements, keyed by a
//          (seed, offset) pair that lives in DEVICE memory (so a captured graph draws fresh numbers at every replay);
//   mask   mask = 1 - maxpool3x3(seeds) (stride 1, padding 1: reference :333-334), its sum accumulated in a device counter;
//          this kernel also advances the RNG offset;
//   apply  y = x * mask * numel / sum(mask) in the reference's operation order (:341) -- also the backward (dy in, dx out).
// The tensors may be NCHW-contiguous or channels_last: seeds / mask share x's strides, the random number of an element depends
// on its LOGICAL index only, so both layouts see the same mask.
#include "common.cuh"

namespace ppy {
namespace {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
// Philox4x32-10: counter (call index, offset), key = seed
__device__ __forceinline__ void philox4x32_10(unsigned long long call, unsigned long long offset, unsigned long long seed, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)call, (uint32_t)(call >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) out[i] = c[i];
}

struct Shape4 { int n, c, h, w; long long sn, sc, sh, sw; };

__device__ __forceinline__ long long addr_of(const Shape4& s, long long i) {     // logical NCHW index -> storage offset
  const int x = (int)(i % s.w); i /= s.w;
  const int y = (int)(i % s.h); i /= s.h;
  const int ch = (int)(i % s.c);
  const long long img = i / s.c;
  return img * s.sn + ch * s.sc + y * s.sh + x * s.sw;
}

__global__ void dropblock_seeds_kernel(uint8_t* __restrict__ seeds, Shape4 s, float gamma, const unsigned long long* __restrict__ rng,
                                       unsigned int* __restrict__ count) {
  const long long total = (long long)s.n * s.c * s.h * s.w;
  const unsigned long long seed = rng[0], offset = rng[1];
  if (blockIdx.x == 0 && threadIdx.x == 0) *count = 0u;
  if (s.sc == 1 && s.c > 1) {
    // channels_last storage: walk the STORAGE order (channel fastest) so the byte stores coalesce; every element still takes
    // word (i & 3) of the Philox call of its logical index i / 4 -- four times the RNG work, the same numbers
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (long long)gridDim.x * blockDim.x) {
      long long t = j;
      const int ch = (int)(t % s.c); t /= s.c;
      const int x = (int)(t % s.w); t /= s.w;
      const int y = (int)(t % s.h);
      const long long img = t / s.h;
      const long long i = ((img * s.c + ch) * s.h + y) * s.w + x;
      uint32_t r[4];
      philox4x32_10((unsigned long long)(i >> 2), offset, seed, r);
      const float u = (float)(r[i & 3] >> 8) * (1.0f / 16777216.0f);
      seeds[img * s.sn + ch * s.sc + y * s.sh + x * s.sw] = u < gamma ? 1 : 0;
    }
    return;
  }
  for (long long call = (long long)blockIdx.x * blockDim.x + threadIdx.x; call * 4 < total; call += (long long)gridDim.x * blockDim.x) {
    uint32_t r[4];
    philox4x32_10((unsigned long long)call, offset, seed, r);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long i = call * 4 + k;
      if (i < total) {
        const float u = (float)(r[k] >> 8) * (1.0f / 16777216.0f);        // uniform in [0, 1), 24 bits like torch.rand(float32)
        seeds[addr_of(s, i)] = u < gamma ? 1 : 0;
      }
    }
  }
}

__global__ void dropblock_mask_kernel(const uint8_t* __restrict__ seeds, uint8_t* __restrict__ mask, Shape4 s, unsigned int* __restrict__ count,
                                      unsigned long long* __restrict__ rng) {
  const long long total = (long long)s.n * s.c * s.h * s.w;
  unsigned int kept = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    int x, y, ch;
    if (s.sc == 1 && s.c > 1) {          // channels_last storage: channel fastest, so a warp's byte loads / stores coalesce
      ch = (int)(t % s.c); t /= s.c;
      x = (int)(t % s.w); t /= s.w;
      y = (int)(t % s.h); t /= s.h;
    } else {
      x = (int)(t % s.w); t /= s.w;
      y = (int)(t % s.h); t /= s.h;
      ch = (int)(t % s.c); t /= s.c;
    }
    const long long base = t * s.sn + ch * s.sc;
    int hit = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= s.h) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= s.w) continue;
        hit |= seeds[base + yy * s.sh + xx * s.sw];
      }
    }
    mask[base + y * s.sh + x * s.sw] = hit ? 0 : 1;
    kept += hit ? 0u : 1u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, o);
  if ((threadIdx.x & 31) == 0 && kept) atomicAdd(count, kept);
  if (blockIdx.x == 0 && threadIdx.x == 0) rng[1] += 1ull;                 // the next draw uses a fresh Philox stream
}

template <typename T>
__global__ void dropblock_apply_kernel(const T* __restrict__ x, T* __restrict__ y, const uint8_t* __restrict__ mask,
                                       const unsigned int* __restrict__ count, long long numel) {
  const float nf = (float)numel, sum = (float)(*count);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (long long)gridDim.x * blockDim.x) {
    // input * mask * elem_numel / elem_sum, left to right (reference :341)
    const float v = __fdiv_rn(__fmul_rn(__fmul_rn(to_f<T>(x[i]), (float)mask[i]), nf), sum);
    y[i] = from_f<T>(v);
  }
}

inline unsigned grid_for(long long work, int threads) {
  long long b = ceil_div(work, threads);
  const long long cap = 148ll * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace
}  // namespace ppy

extern "C" {
using namespace ppy;

/* strides: element strides of the logical (n, c, h, w) dimensions, shared by seeds and mask (both uint8). */
int ppy_dropblock_mask(uint8_t* seeds, uint8_t* mask, int n, int c, int h, int w, long long sn, long long sc, long long sh, long long sw,
                       int block_size, float gamma, unsigned long long* rng_state, unsigned int* count, ppy_stream_t s) {
  PPY_REQUIRE(seeds && mask && rng_state && count && n > 0 && c > 0 && h > 0 && w > 0);
  if (block_size != 3) return PPY_ERR_UNSUPPORTED;      // the reference's max_pool2d(block_size, padding=1) keeps the shape for 3 only
  const Shape4 sp = {n, c, h, w, sn, sc, sh, sw};
  const long long total = (long long)n * c * h * w;
  dropblock_seeds_kernel<<<grid_for(ceil_div(total, 4), 256), 256, 0, as_stream(s)>>>(seeds, sp, gamma, rng_state, count);
  int rc = check_launch();
  if (rc) return rc;
  dropblock_mask_kernel<<<grid_for(total, 256), 256, 0, as_stream(s)>>>(seeds, mask, sp, count, rng_state);
  return check_launch();
}

/* Only the second stage, from caller-provided seeds (tests inject the seed matrix of a torch.rand draw). */
int ppy_dropblock_mask_from_seeds(const uint8_t* seeds, uint8_t* mask, int n, int c, int h, int w, long long sn, long long sc, long long sh,
                                  long long sw, unsigned long long* rng_state, unsigned int* count, ppy_stream_t s) {
  PPY_REQUIRE(seeds && mask && rng_state && count && n > 0 && c > 0 && h > 0 && w > 0);
  const Shape4 sp = {n, c, h, w, sn, sc, sh, sw};
  if (check_cuda(cudaMemsetAsync(count, 0, sizeof(unsigned int), as_stream(s)))) return PPY_ERR_CUDA;
  dropblock_mask_kernel<<<grid_for((long long)n * c * h * w, 256), 256, 0, as_stream(s)>>>(seeds, mask, sp, count, rng_state);
  return check_launch();
}

int ppy_dropblock_apply(const void* x, void* y, const uint8_t* mask, const unsigned int* count, long long numel, int dtype, ppy_stream_t s) {
  PPY_REQUIRE(x && y && mask && count && numel > 0);
  if (dtype == PPY_F32) dropblock_apply_kernel<float><<<grid_for(numel, 256), 256, 0, as_stream(s)>>>((const float*)x, (float*)y, mask, count, numel);
  else if (dtype == PPY_BF16) dropblock_apply_kernel<__nv_bfloat16><<<grid_for(numel, 256), 256, 0, as_stream(s)>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, mask, count, numel);
  else return PPY_ERR_INVALID;
  return check_launch();
}

}  // extern "C"
