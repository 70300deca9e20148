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

// channels_last storage with c % 16 == 0: one thread = 16 consecutive channels of one pixel (one 16-byte store / 9 16-byte loads
// instead of 16 byte-wide accesses with their index divisions); every element still takes word (i & 3) of the Philox call of its
// logical NCHW index i >> 2, so the numbers are those of the generic kernels.
__global__ void __launch_bounds__(256) dropblock_seeds_cl16_kernel(uint8_t* __restrict__ seeds, Shape4 s, float gamma,
                                                                   const unsigned long long* __restrict__ rng, unsigned int* __restrict__ count) {
  const unsigned long long seed = rng[0], offset = rng[1];
  if (blockIdx.x == 0 && threadIdx.x == 0) *count = 0u;
  const int cg = s.c / 16;
  const long long groups = (long long)s.n * s.h * s.w * cg, plane = (long long)s.h * s.w;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < groups; j += (long long)gridDim.x * blockDim.x) {
    long long t = j;
    const int g = (int)(t % cg); t /= cg;
    const int x = (int)(t % s.w); t /= s.w;
    const int y = (int)(t % s.h);
    const long long img = t / s.h;
    long long i = ((img * s.c + g * 16) * s.h + y) * s.w + x;
    uint32_t words[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int k = 0; k < 16; ++k, i += plane) {
      uint32_t r[4];
      philox4x32_10((unsigned long long)(i >> 2), offset, seed, r);
      const int sel = (int)(i & 3);
      const uint32_t rk = sel == 0 ? r[0] : sel == 1 ? r[1] : sel == 2 ? r[2] : r[3];
      const float u = (float)(rk >> 8) * (1.0f / 16777216.0f);
      if (u < gamma) words[k >> 2] |= 1u << (8 * (k & 3));
    }
    *reinterpret_cast<uint4*>(seeds + img * s.sn + y * s.sh + x * s.sw + g * 16) = make_uint4(words[0], words[1], words[2], words[3]);
  }
}

__global__ void __launch_bounds__(256) dropblock_mask_cl16_kernel(const uint8_t* __restrict__ seeds, uint8_t* __restrict__ mask, Shape4 s,
                                                                  unsigned int* __restrict__ count, unsigned long long* __restrict__ rng) {
  const int cg = s.c / 16;
  const long long groups = (long long)s.n * s.h * s.w * cg;
  unsigned int kept = 0;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < groups; j += (long long)gridDim.x * blockDim.x) {
    long long t = j;
    const int g = (int)(t % cg); t /= cg;
    const int x = (int)(t % s.w); t /= s.w;
    const int y = (int)(t % s.h);
    const long long base = (t / s.h) * s.sn + g * 16;
    uint4 hit = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= s.h) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= s.w) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(seeds + base + yy * s.sh + xx * s.sw));
        hit.x |= v.x; hit.y |= v.y; hit.z |= v.z; hit.w |= v.w;
      }
    }
    const uint4 m = make_uint4(~hit.x & 0x01010101u, ~hit.y & 0x01010101u, ~hit.z & 0x01010101u, ~hit.w & 0x01010101u);   // seeds are 0 / 1 bytes
    *reinterpret_cast<uint4*>(mask + base + y * s.sh + x * s.sw) = m;
    kept += __popc(m.x) + __popc(m.y) + __popc(m.z) + __popc(m.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, o);
  if ((threadIdx.x & 31) == 0 && kept) atomicAdd(count, kept);
  if (blockIdx.x == 0 && threadIdx.x == 0) rng[1] += 1ull;
}

inline bool cl16_ok(const void* a, const void* b, const Shape4& s) {
  return s.sc == 1 && s.c % 16 == 0 && s.sw % 16 == 0 && s.sh % 16 == 0 && s.sn % 16 == 0 &&
         ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
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

// eight elements per thread (16-byte loads of bf16, 8 mask bytes); same arithmetic
template <typename T>
__global__ void __launch_bounds__(256) dropblock_apply8_kernel(const T* __restrict__ x, T* __restrict__ y, const uint8_t* __restrict__ mask,
                                                               const unsigned int* __restrict__ count, long long numel) {
  const float nf = (float)numel, sum = (float)(*count);
  const long long groups = numel / 8;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < groups; j += (long long)gridDim.x * blockDim.x) {
    __align__(16) T in[8];
    __align__(16) T out[8];
    constexpr int NV = (int)(sizeof(T) * 8 / 16);
#pragma unroll
    for (int q = 0; q < NV; ++q) reinterpret_cast<uint4*>(in)[q] = __ldg(reinterpret_cast<const uint4*>(x + j * 8) + q);
    const uint2 mk = __ldg(reinterpret_cast<const uint2*>(mask + j * 8));
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float m = (float)(((k < 4 ? mk.x : mk.y) >> (8 * (k & 3))) & 0xffu);
      out[k] = from_f<T>(__fdiv_rn(__fmul_rn(__fmul_rn(to_f<T>(in[k]), m), nf), sum));
    }
#pragma unroll
    for (int q = 0; q < NV; ++q) reinterpret_cast<uint4*>(y + j * 8)[q] = reinterpret_cast<const uint4*>(out)[q];
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
  if (cl16_ok(seeds, mask, sp)) {
    dropblock_seeds_cl16_kernel<<<grid_for(total / 16, 256), 256, 0, as_stream(s)>>>(seeds, sp, gamma, rng_state, count);
    int rc = check_launch();
    if (rc) return rc;
    dropblock_mask_cl16_kernel<<<grid_for(total / 16, 256), 256, 0, as_stream(s)>>>(seeds, mask, sp, count, rng_state);
    return check_launch();
  }
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
  if (cl16_ok(seeds, mask, sp)) dropblock_mask_cl16_kernel<<<grid_for((long long)n * c * h * w / 16, 256), 256, 0, as_stream(s)>>>(seeds, mask, sp, count, rng_state);
  else dropblock_mask_kernel<<<grid_for((long long)n * c * h * w, 256), 256, 0, as_stream(s)>>>(seeds, mask, sp, count, rng_state);
  return check_launch();
}

int ppy_dropblock_apply(const void* x, void* y, const uint8_t* mask, const unsigned int* count, long long numel, int dtype, ppy_stream_t s) {
  PPY_REQUIRE(x && y && mask && count && numel > 0);
  const bool vec8 = numel % 8 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0 && (reinterpret_cast<uintptr_t>(mask) & 7) == 0;
  if (vec8 && dtype == PPY_F32) dropblock_apply8_kernel<float><<<grid_for(numel / 8, 256), 256, 0, as_stream(s)>>>((const float*)x, (float*)y, mask, count, numel);
  else if (vec8 && dtype == PPY_BF16) dropblock_apply8_kernel<__nv_bfloat16><<<grid_for(numel / 8, 256), 256, 0, as_stream(s)>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, mask, count, numel);
  else if (dtype == PPY_F32) dropblock_apply_kernel<float><<<grid_for(numel, 256), 256, 0, as_stream(s)>>>((const float*)x, (float*)y, mask, count, numel);
  else if (dtype == PPY_BF16) dropblock_apply_kernel<__nv_bfloat16><<<grid_for(numel, 256), 256, 0, as_stream(s)>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, mask, count, numel);
  else return PPY_ERR_INVALID;
  return check_launch();
}

}  // extern "C"
