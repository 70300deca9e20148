// First layer of the ResNet-vd deep stem fused with the input layout change: reads the NCHW fp32 image batch
// exactly as Decode.predict uploads it (reference model/decode_np.py:142-147), applies conv 3x3 / stride 2 / pad 1
// (3 -> 32, reference model/resnet_vd.py:100) + folded BN + activation in fp32 SIMT math and writes NHWC
// (bf16 or fp32).  K = 27 is far too small for the tensor pipe and the layer is HBM-bound (142 MB in, 189 MB out
// at bs=32 x 608^2); the weights ride in the kernel parameter (constant bank), so every FFMA takes its weight as a
// constant operand and each thread keeps its 32 output channels of one pixel in registers.
//
// bf16 output (the tensor-core path of the engine) runs stem_umma_kernel instead: the SIMT kernel is bound by fp32
// issue (864 FFMA per output pixel, 2.1 TB/s of its 331 MB), not by HBM.  There each thread stages the 27 taps of ONE
// output pixel as a bf16 row of a [128 pixels x K=32] SWIZZLE_128B operand tile, one elected thread issues two
// tcgen05.mma (128x32x16) against the [32 x 32] weight tile into a 32-column TMEM accumulator, and the same thread
// reads its pixel's 32 channels back (tcgen05.ld 32x32b.x32) for scale/shift/activation and four 16-byte stores.
// Several CTAs per SM (20 KB of shared memory, 32 TMEM columns each) overlap each other's load / MMA / store phases.
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include "common.cuh"

namespace ppy {
namespace {

constexpr int STEM_CIN = 3, STEM_COUT = 32, STEM_TAPS = 27;

struct StemParams {
  float w[STEM_TAPS][STEM_COUT];   // [c*9 + ky*3 + kx][cout]
  float scale[STEM_COUT];
  float shift[STEM_COUT];
};

// U8: x is the resized uint8 RGB image batch [n, h, w, 3] (HWC, what cv2 hands Decode.process_image after ResizeImage) and
// `lut` [3][256] maps a byte of channel c to the reference's normalised float ((u / 255 - mean[c]) / std[c], computed on the
// host with the reference's numpy expression, so NormalizeImage + Permute are bit-exact by construction): a quarter of the
// upload bytes, no separate normalisation pass.
template <typename T, bool U8 = false>
__global__ void __launch_bounds__(128) stem_conv_kernel(const float* __restrict__ x, int n, int h, int w, int ho, int wo,
                                                        const __grid_constant__ StemParams prm, float slope, T* __restrict__ y,
                                                        int y_ld, long long y_plane = 0 /* > 0: T = __half, fp16 hi/lo pair output */,
                                                        const float* __restrict__ lut = nullptr) {
  __shared__ float s_lut[U8 ? 3 * 256 : 1];
  if (U8) {
    for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) s_lut[U8 ? i : 0] = __ldg(lut + i);
    __syncthreads();
  }
  const uint8_t* x8 = reinterpret_cast<const uint8_t*>(x);
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
          if (iy >= 0 && iy < h && ix >= 0 && ix < w) {
            if (U8) v = s_lut[U8 ? c * 256 + (int)__ldg(x8 + (((long long)img * h + iy) * w + ix) * 3 + c) : 0];
            else v = __ldg(xi + ((long long)c * h + iy) * w + ix);
          }
#pragma unroll
          for (int co = 0; co < STEM_COUT; ++co) acc[co] = fmaf(v, prm.w[c * 9 + ky * 3 + kx][co], acc[co]);
        }
      }
    }
    T* dst = y + i * y_ld;
    if (y_plane > 0) {                             // PPY_F16X2: split every fp32 result into the hi and the lo plane
#pragma unroll
      for (int co = 0; co < STEM_COUT; co += 8) {
        uint4 hi, lo;
        __half2* hp = reinterpret_cast<__half2*>(&hi);
        __half2* lp = reinterpret_cast<__half2*>(&lo);
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
          float f0 = acc[co + k] * prm.scale[co + k] + prm.shift[co + k], f1 = acc[co + k + 1] * prm.scale[co + k + 1] + prm.shift[co + k + 1];
          f0 = f0 > 0.f ? f0 : f0 * slope;
          f1 = f1 > 0.f ? f1 : f1 * slope;
          hp[k >> 1] = __floats2half2_rn(f0, f1);
          const float2 hf = __half22float2(hp[k >> 1]);
          lp[k >> 1] = __floats2half2_rn(f0 - hf.x, f1 - hf.y);
        }
        *reinterpret_cast<uint4*>(dst + co) = hi;
        *reinterpret_cast<uint4*>(dst + y_plane + co) = lo;
      }
      continue;
    }
#pragma unroll
    for (int co = 0; co < STEM_COUT; co += 16 / sizeof(T)) {
      uint4 raw;
      T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
      for (int k = 0; k < (int)(16 / sizeof(T)); ++k) {
        float f = acc[co + k] * prm.scale[co + k] + prm.shift[co + k];
        f = f > 0.f ? f : f * slope;
        if constexpr (!std::is_same<T, __half>::value) e[k] = from_f<T>(f);
      }
      *reinterpret_cast<uint4*>(dst + co) = raw;
    }
  }
}


// ---------------------------------------------------------------------------------------------
// tcgen05 path (bf16 output)
// ---------------------------------------------------------------------------------------------
struct StemUmmaParams {
  uint32_t w[STEM_COUT][16];       // bf16 pairs, K order k = c*9 + ky*3 + kx, k = 27..31 zero
  float scale[STEM_COUT];
  float shift[STEM_COUT];
};

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
// K-major SWIZZLE_128B operand descriptor (rows 128 B apart, 8-row groups 1024 B apart) and the bf16 x bf16 -> fp32
// instruction descriptor: the same encodings as conv_umma.cu
__device__ __forceinline__ uint64_t stem_smem_desc(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
constexpr uint32_t kStemIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(STEM_COUT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
// Tile = 2 output rows x 64 output pixels (128 UMMA rows: warp w -> row w >> 1, pixels (w & 1) * 32 + lane).  Its input
// patch (3 channels x 5 rows x 132 floats, columns 2*ox0 - 4 .. 2*ox0 + 127, 16-byte aligned because w % 4 == 0) arrives by
// 16-byte cp.async with zero fill outside the image, double-buffered so the next tile's patch is in flight during this
// tile's MMA / epilogue; each warp's 32 output pixels (2 KB, contiguous in NHWC with ld = 32) are transposed through
// shared memory into four fully coalesced 512-byte store instructions.  No scalar global loads, no strided stores.
constexpr int kPatchW = 132, kPatchRows = 5;                    // 3 x 5 x 132 = 1980 floats per patch
constexpr int kPatchBytes = 7936;                              // 1980 * 4 rounded up to 128
constexpr int kPatchVecs = STEM_CIN * kPatchRows * (kPatchW / 4);                              // 495 x 16 bytes
constexpr int kStemSmem = 128 * 128 + STEM_COUT * 128 + 2 * kPatchBytes + 128 * 64 + 1024;     // A, W, 2 patches, out staging, slack

__global__ void __launch_bounds__(128) stem_umma_kernel(const float* __restrict__ x, int h, int w, int ho, int wo,
                                                        const __grid_constant__ StemUmmaParams prm, float slope,
                                                        __nv_bfloat16* __restrict__ y, int tiles_x, int tiles_y, int num_tiles) {
  extern __shared__ uint8_t stem_smem_raw[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_addr = (s_u32(stem_smem_raw) + 1023u) & ~1023u;
  const uint32_t b_addr = a_addr + 128 * 128;
  const uint32_t patch_addr = b_addr + STEM_COUT * 128;
  const uint32_t out_addr = patch_addr + 2 * kPatchBytes + (uint32_t)warp * 2048u;   // this warp's 32 pixels x 64 bytes
  const uint32_t bar = s_u32(&s_bar);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&s_tmem)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {  // weight tile: row = cout, 4 chunks of 16 bytes (K = 32 bf16), chunk j of row r at ((j ^ (r & 7)) << 4)
    const int r = tid >> 2, j = tid & 3;
    const uint32_t dst = b_addr + (uint32_t)r * 128u + (((uint32_t)j ^ (uint32_t)(r & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(prm.w[r][4 * j]), "r"(prm.w[r][4 * j + 1]),
                 "r"(prm.w[r][4 * j + 2]), "r"(prm.w[r][4 * j + 3]) : "memory");
  }
  // Each CTA walks a contiguous run of tiles, so (tx, ty, img) advance by increments (no divisions in the loop), and the
  // patch vectors a thread copies -- item = tid + 128 k -> (channel, patch row, 16-byte column) -- are the same for every
  // tile: their shared-memory offsets and image-relative source offsets are computed once.
  const int per_cta = (num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int tile_begin = (int)blockIdx.x * per_cta;
  const int tile_end = tile_begin + per_cta < num_tiles ? tile_begin + per_cta : num_tiles;
  struct Pos { int tx, ty, img; };
  auto advance = [&](Pos& p) {
    if (++p.tx == tiles_x) { p.tx = 0; if (++p.ty == tiles_y) { p.ty = 0; ++p.img; } }
  };
  uint32_t it_dst[4];
  int it_rel[4], it_pr[4], it_col[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int item = tid + 128 * k;
    const int c = item / (kPatchRows * 33), rem = item - c * (kPatchRows * 33), pr = rem / 33, q = rem - pr * 33;
    it_dst[k] = patch_addr + (uint32_t)((c * kPatchRows + pr) * kPatchW + 4 * q) * 4u;
    it_rel[k] = (c * h + pr) * w + 4 * q;
    it_pr[k] = item < kPatchVecs ? pr : (1 << 28);        // past the end: never valid
    it_col[k] = 4 * q;
  }
  // 16-byte cp.async of one tile's patch; out-of-image vectors are zero-filled (src-size 0, address clamped to x)
  auto fetch = [&](const Pos& p, int buf) {
    const int iy_first = p.ty * 4 - 1, col_first = p.tx * 128 - 4;
    const float* base = x + ((size_t)p.img * (size_t)(STEM_CIN * h * w) + (size_t)((long long)iy_first * w + col_first));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k == 3 && tid + 384 >= kPatchVecs) break;
      const bool ok = (unsigned)(iy_first + it_pr[k]) < (unsigned)h && (unsigned)(col_first + it_col[k]) < (unsigned)w;
      const float* src = ok ? base + it_rel[k] : x;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(it_dst[k] + (uint32_t)buf * kPatchBytes), "l"(src),
                   "r"(ok ? 16u : 0u) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  Pos cur, nxt;
  cur.tx = tile_begin % tiles_x;
  cur.ty = (tile_begin / tiles_x) % tiles_y;
  cur.img = tile_begin / (tiles_x * tiles_y);
  nxt = cur;
  if (tile_begin < tile_end) fetch(cur, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  const uint32_t a_row = a_addr + (uint32_t)tid * 128u, sw = (uint32_t)(tid & 7);
  const int prow = warp >> 1, ppx = (warp & 1) * 32 + lane;          // this thread's pixel inside the tile
  uint32_t phase = 0;
  int buf = 0;
  for (int tile = tile_begin; tile < tile_end; ++tile, buf ^= 1, cur = nxt) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                                   // patch[buf] landed for every thread
    const float* patch = reinterpret_cast<const float*>(stem_smem_raw + (patch_addr - s_u32(stem_smem_raw)) + buf * kPatchBytes);
    float v[28];
#pragma unroll
    for (int c = 0; c < STEM_CIN; ++c)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
          v[c * 9 + ky * 3 + kx] = patch[(c * kPatchRows + 2 * prow + ky) * kPatchW + 2 * ppx + 3 + kx];
    v[27] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t q[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = j * 8 + e * 2;
        q[e] = k < 28 ? pack2(v[k], v[k + 1]) : 0u;
      }
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_row + (((uint32_t)j ^ sw) << 4)), "r"(q[0]), "r"(q[1]),
                   "r"(q[2]), "r"(q[3]) : "memory");
    }
    // generic-proxy writes -> visible to the tensor core's async proxy; every thread's previous tcgen05.ld has completed
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
      uint32_t elected;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
      if (elected) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const uint64_t da = stem_smem_desc(a_addr + k * 32), db = stem_smem_desc(b_addr + k * 32);
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tmem), "l"(da), "l"(db), "r"(kStemIdesc), "r"((uint32_t)k) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
      }
      __syncwarp();
    }
    // next tile's patch: issued only now -- a cp.async in flight ahead of fence.proxy.async would be waited for by the fence
    advance(nxt);
    if (tile + 1 < tile_end) fetch(nxt, buf ^ 1);
    {
      uint32_t done;
      do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(phase) : "memory");
      } while (!done);
    }
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t acc[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]), "=r"(acc[7]),
          "=r"(acc[8]), "=r"(acc[9]), "=r"(acc[10]), "=r"(acc[11]), "=r"(acc[12]), "=r"(acc[13]), "=r"(acc[14]), "=r"(acc[15]),
          "=r"(acc[16]), "=r"(acc[17]), "=r"(acc[18]), "=r"(acc[19]), "=r"(acc[20]), "=r"(acc[21]), "=r"(acc[22]), "=r"(acc[23]),
          "=r"(acc[24]), "=r"(acc[25]), "=r"(acc[26]), "=r"(acc[27]), "=r"(acc[28]), "=r"(acc[29]), "=r"(acc[30]), "=r"(acc[31])
        : "r"(tmem + ((uint32_t)(warp * 32) << 16)) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int co = 0; co < STEM_COUT; co += 8) {
      uint32_t q[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float f0 = __uint_as_float(acc[co + 2 * e]) * prm.scale[co + 2 * e] + prm.shift[co + 2 * e];
        const float f1 = __uint_as_float(acc[co + 2 * e + 1]) * prm.scale[co + 2 * e + 1] + prm.shift[co + 2 * e + 1];
        q[e] = pack2(fmaxf(f0, f0 * slope), fmaxf(f1, f1 * slope));   // slope in [0, 1]: relu / leaky / identity
      }
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(out_addr + (uint32_t)lane * 64u + (uint32_t)co * 2u), "r"(q[0]),
                   "r"(q[1]), "r"(q[2]), "r"(q[3]) : "memory");
    }
    __syncwarp();
    {  // the warp's 32 pixels x 64 bytes are contiguous in NHWC (ld = 32): four fully coalesced 512-byte store instructions
      const int img = cur.img, oy = cur.ty * 2 + prow, ox = cur.tx * 64 + (warp & 1) * 32;
      int live = wo - ox;
      live = oy < ho ? (live > 32 ? 32 : live) : 0;
      uint8_t* dst = reinterpret_cast<uint8_t*>(y + ((size_t)((size_t)img * ho + oy) * wo + ox) * STEM_COUT);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int off = k * 512 + lane * 16;
        uint4 q;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(out_addr + (uint32_t)off));
        if (off < live * 64) *reinterpret_cast<uint4*>(dst + off) = q;
      }
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
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
  const float slope = act == PPY_ACT_RELU ? 0.f : (act == PPY_ACT_LEAKY ? 0.1f : 1.f);
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  const long long total = (long long)n * ho * wo;
  static const bool no_umma = getenv("PPY_NO_STEM_UMMA") != nullptr;   // debug knob: bf16 output through the SIMT kernel
  if (y_dtype == PPY_BF16 && !no_umma && ppy_conv_bf16_supported() && w % 4 == 0 && y_ld == STEM_COUT &&
      (reinterpret_cast<uintptr_t>(x_nchw) & 15) == 0) {
    const int tiles_x = (wo + 63) / 64, tiles_y = (ho + 1) / 2;
    PPY_REQUIRE((long long)n * tiles_x * tiles_y < 0x7FFFFFFFll && (long long)STEM_CIN * h * w < 0x7FFFFFFFll);
    StemUmmaParams up;
    memset(&up, 0, sizeof(up));
    for (int co = 0; co < STEM_COUT; ++co) {
      for (int t = 0; t < STEM_TAPS; ++t) {
        const __nv_bfloat16 b = __float2bfloat16_rn(weight_oihw_host[co * STEM_TAPS + t]);
        up.w[co][t >> 1] |= (uint32_t)(*reinterpret_cast<const uint16_t*>(&b)) << ((t & 1) * 16);
      }
      up.scale[co] = scale_host[co];
      up.shift[co] = shift_host[co];
    }
    static DeviceOnce attr_once;
    if (attr_once.first()) {
      if (check_cuda(cudaFuncSetAttribute(stem_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStemSmem))) return PPY_ERR_CUDA;
    }
    const int tiles = n * tiles_x * tiles_y;
    static const int per_sm = getenv("PPY_STEM_CTAS") ? atoi(getenv("PPY_STEM_CTAS")) : 4;   // measured: 3 -> 96.6, 4 -> 90.5, 5 -> 110.2 us (bs 32 x 608^2)
    const int grid = tiles < 148 * per_sm ? tiles : 148 * per_sm;
    stem_umma_kernel<<<grid, 128, kStemSmem, as_stream(s)>>>(x_nchw, h, w, ho, wo, up, slope, (__nv_bfloat16*)y, tiles_x, tiles_y, tiles);
    return check_launch();
  }
  StemParams prm;
  for (int co = 0; co < STEM_COUT; ++co) {
    for (int t = 0; t < STEM_TAPS; ++t) prm.w[t][co] = weight_oihw_host[co * STEM_TAPS + t];   // OIHW: [co][c][ky][kx]
    prm.scale[co] = scale_host[co];
    prm.shift[co] = shift_host[co];
  }
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

// Same layer reading the resized uint8 HWC batch through the normalisation table (see stem_conv_kernel<T, U8>): y_dtype PPY_F32,
// PPY_BF16 (fp32 SIMT math) or PPY_F16X2 (y_plane > 0).  lut: DEVICE pointer, [3][256] floats.
extern "C" int ppy_stem_conv3x3s2_u8(const uint8_t* x_nhwc_u8, int n, int h, int w, const float* lut, const float* weight_oihw_host,
                                     const float* scale_host, const float* shift_host, int cout, int act, void* y, int y_ld,
                                     int y_dtype, long long y_plane, ppy_stream_t s) {
  using namespace ppy;
  PPY_REQUIRE(x_nhwc_u8 && lut && weight_oihw_host && scale_host && shift_host && y);
  PPY_REQUIRE(n > 0 && h > 1 && w > 1 && cout == STEM_COUT && y_ld >= cout);
  PPY_REQUIRE(act == PPY_ACT_NONE || act == PPY_ACT_RELU || act == PPY_ACT_LEAKY);
  PPY_REQUIRE((reinterpret_cast<uintptr_t>(y) & 15) == 0 && (y_ld * dtype_size(y_dtype)) % 16 == 0);
  PPY_REQUIRE(y_dtype != PPY_F16X2 || (y_plane > 0 && y_plane % 8 == 0));
  const float slope = act == PPY_ACT_RELU ? 0.f : (act == PPY_ACT_LEAKY ? 0.1f : 1.f);
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  StemParams prm;
  for (int co = 0; co < STEM_COUT; ++co) {
    for (int t = 0; t < STEM_TAPS; ++t) prm.w[t][co] = weight_oihw_host[co * STEM_TAPS + t];
    prm.scale[co] = scale_host[co];
    prm.shift[co] = shift_host[co];
  }
  long long blocks = ceil_div((long long)n * ho * wo, 128);
  if (blocks > 148 * 64) blocks = 148 * 64;
  const float* xf = reinterpret_cast<const float*>(x_nhwc_u8);
  cudaStream_t st = as_stream(s);
  if (y_dtype == PPY_F16X2) stem_conv_kernel<__half, true><<<(unsigned)blocks, 128, 0, st>>>(xf, n, h, w, ho, wo, prm, slope, (__half*)y, y_ld, y_plane, lut);
  else if (y_dtype == PPY_BF16) stem_conv_kernel<__nv_bfloat16, true><<<(unsigned)blocks, 128, 0, st>>>(xf, n, h, w, ho, wo, prm, slope, (__nv_bfloat16*)y, y_ld, 0, lut);
  else if (y_dtype == PPY_F32) stem_conv_kernel<float, true><<<(unsigned)blocks, 128, 0, st>>>(xf, n, h, w, ho, wo, prm, slope, (float*)y, y_ld, 0, lut);
  else return PPY_ERR_INVALID;
  return check_launch();
}

// Same layer with a PPY_F16X2 output (fp32 SIMT math, results split into fp16 hi/lo planes): first kernel of the fp32-grade
// tensor-core engine.
extern "C" int ppy_stem_conv3x3s2_f16x2(const float* x_nchw, int n, int h, int w, const float* weight_oihw_host,
                                        const float* scale_host, const float* shift_host, int cout, int act, void* y, int y_ld,
                                        long long y_plane, ppy_stream_t s) {
  using namespace ppy;
  PPY_REQUIRE(x_nchw && weight_oihw_host && scale_host && shift_host && y);
  PPY_REQUIRE(n > 0 && h > 1 && w > 1 && cout == STEM_COUT && y_ld >= cout && y_plane > 0 && y_plane % 8 == 0);
  PPY_REQUIRE(act == PPY_ACT_NONE || act == PPY_ACT_RELU || act == PPY_ACT_LEAKY);
  PPY_REQUIRE((reinterpret_cast<uintptr_t>(y) & 15) == 0 && (y_ld * 2) % 16 == 0);
  const float slope = act == PPY_ACT_RELU ? 0.f : (act == PPY_ACT_LEAKY ? 0.1f : 1.f);
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  StemParams prm;
  for (int co = 0; co < STEM_COUT; ++co) {
    for (int t = 0; t < STEM_TAPS; ++t) prm.w[t][co] = weight_oihw_host[co * STEM_TAPS + t];
    prm.scale[co] = scale_host[co];
    prm.shift[co] = shift_host[co];
  }
  long long blocks = ceil_div((long long)n * ho * wo, 128);
  if (blocks > 148 * 64) blocks = 148 * 64;
  stem_conv_kernel<__half><<<(unsigned)blocks, 128, 0, as_stream(s)>>>(x_nchw, n, h, w, ho, wo, prm, slope, (__half*)y, y_ld, y_plane);
  return check_launch();
}
