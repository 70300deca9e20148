// bf16 implicit-GEMM convolution on the sm_100a tensor cores (tcgen05.mma, fp32 accumulators in TMEM),
// with fused scale/shift (+CoordConv bias map, +residual), activation and optional 2x nearest upsample
// in the epilogue.  Also the fused DCNv2 kernel: the same GEMM pipeline with a producer that bilinearly
// samples the NHWC input at the learned offsets and writes the modulated bf16 A tile straight into the
// swizzled shared-memory stage -- no im2col / gather temporaries in HBM (reference
// model/custom_layers.py:571-676 materialises ~1.3 GB of them per layer at bs=32).
//
//   GEMM view: D[M x N] = A[M x K] * B[N x K]^T,  M = n*ho*wo pixels, N = cout, K = kh*kw*cin,
//   K index = (ky*kw + kx)*cin + c.  CTA tile 128 x BLOCK_N x 64, one tile per CTA.
//
//   warps 0-3  A producers: thread r owns tile row r; per 64-wide K block it issues eight 16-byte
//              cp.async (zero-fill outside the image) into the 128B-swizzled K-major stage; later the
//              epilogue: tcgen05.ld 32 columns at a time -> scale/shift/residual/act -> 64/128-byte stores
//   warp 4     B producer: one thread issues TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) of the packed
//              weight tile [BLOCK_N x 64] with mbarrier complete_tx
//   warp 5     TMEM allocator + MMA issuer: one thread issues 4 x tcgen05.mma (K=16 each) per stage and
//              tcgen05.commit's the stage back to the producers
//
// Shared-memory operand layout is the canonical K-major SWIZZLE_128B one: row r of a stage lives at
// r*128 bytes, its 16-byte chunk j at ((j ^ (r & 7)) << 4); 8-row groups are 1024 bytes apart (SBO).
#include <cuda.h>
#include "common.cuh"

namespace ppy {

int validate_conv(const ppy_conv_params* p, int elem_bytes, int* ho, int* wo);

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                       // bf16 elements = 128 bytes = one swizzle row
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int NUM_THREADS = 192;
constexpr int CP_LAG = 2;                         // cp.async groups in flight per producer thread

template <int BN> struct TileCfg {
  static constexpr int kStages = (BN == 128) ? 3 : 4;
  static constexpr int kBStageBytes = BN * BLOCK_K * 2;
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;      // power of two for BN in {32,64,128,256}
  static constexpr int kSmemBytes = kStages * (A_STAGE_BYTES + kBStageBytes) + 1024 /*align slack*/ + 256 /*barriers*/;
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4, [16,30) LBO>>4 (unused for swizzled K-major: 1), [32,46) SBO>>4 = 1024>>4,
// [46,48) version = 1 (Blackwell), [61,64) layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) at [4,6), a/b format BF16 (1) at
// [7,10)/[10,13), K-major A and B, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ float bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
template <int BN, bool DCN>
__global__ void __launch_bounds__(NUM_THREADS)
conv_umma_kernel(const ppy_conv_params p, const int ho, const int wo, const int num_kb,
                 const __grid_constant__ CUtensorMap tmap_b) {
  using Cfg = TileCfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + S * A_STAGE_BYTES;
  const uint32_t bars = smem_b + S * Cfg::kBStageBytes;      // full[S], empty[S], tmem_full, tmem_ptr
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_ptr_slot = reinterpret_cast<volatile uint32_t*>(gen_base + S * (A_STAGE_BYTES + Cfg::kBStageBytes) + (2 * S + 1) * 8);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
  const uint32_t tmem_full_bar = bars + 8u * (2 * S);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long M = (long long)p.n * ho * wo;
  const long long m0 = (long long)blockIdx.y * BLOCK_M;
  const int n0 = blockIdx.x * BN;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), BLOCK_M + 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_slot)), Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_slot;

  if (warp < 4) {
    // ======================= A producer (thread = tile row) =======================
    const int r = tid;
    const long long m = m0 + r;
    const bool valid = m < M;
    int img = 0, oy = 0, ox = 0;
    if (valid) { ox = (int)(m % wo); oy = (int)((m / wo) % ho); img = (int)(m / ((long long)wo * ho)); }
    const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(p.x);
    const int taps = p.kh * p.kw;
    const uint32_t row_off = (uint32_t)r * 128u;
    const uint32_t sw = (uint32_t)(r & 7);
    if (!DCN) {
      int tap = 0, c = 0, ky = 0, kx = 0;
      const int iy0 = oy * p.stride - p.pad, ix0 = ox * p.stride - p.pad;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % S;
        mbar_wait(empty_bar(s), ((kb / S) & 1) ^ 1);
        const uint32_t dst_row = smem_a + s * A_STAGE_BYTES + row_off;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int iy = iy0 + ky, ix = ix0 + kx;
          const bool ok = valid && tap < taps && iy >= 0 && iy < p.h && ix >= 0 && ix < p.w;
          const __nv_bfloat16* src = ok ? x + (((long long)img * p.h + iy) * p.w + ix) * p.x_ld + c : x;
          cp_async16(dst_row + (((uint32_t)j ^ sw) << 4), src, ok ? 16u : 0u);
          c += 8;
          if (c >= p.cin) { c = 0; ++tap; if (++kx == p.kw) { kx = 0; ++ky; } }
        }
        cp_async_commit();
        if (kb >= CP_LAG) {
          cp_async_wait<CP_LAG>();
          fence_proxy_async();
          mbar_arrive(full_bar((kb - CP_LAG) % S));
        }
      }
      // drain the last CP_LAG groups
      cp_async_wait<0>();
      fence_proxy_async();
      for (int kb = (num_kb > CP_LAG ? num_kb - CP_LAG : 0); kb < num_kb; ++kb) mbar_arrive(full_bar(kb % S));
    } else {
      // DCNv2: one tap per K block (cin % 64 == 0); bilinear sample + modulation, fp32 math, bf16 store
      const float* om = valid ? p.offset_mask + m * p.om_ld : nullptr;
      const int kb_per_tap = p.cin / BLOCK_K;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % S;
        const int tap = kb / kb_per_tap, c0 = (kb % kb_per_tap) * BLOCK_K;
        float w4[4] = {0.f, 0.f, 0.f, 0.f};
        const __nv_bfloat16* src4[4] = {nullptr, nullptr, nullptr, nullptr};
        if (valid && tap < taps) {
          const int ky = tap / p.kw, kx = tap % p.kw;
          const float dy = __ldg(om + 2 * tap), dx = __ldg(om + 2 * tap + 1), ml = __ldg(om + 2 * taps + tap);
          const float mask = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-ml)));
          const float py = (float)(oy * p.stride - p.pad + ky) + dy, px = (float)(ox * p.stride - p.pad + kx) + dx;
          const float fy = floorf(py), fx = floorf(px);
          const float ly = py - fy, lx = px - fx, hy = 1.f - ly, hx = 1.f - lx;
          const int y0 = (int)fy, x0 = (int)fx;
          const float wq[4] = {hy * hx, hy * lx, ly * hx, ly * lx};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int yy = y0 + (q >> 1), xx = x0 + (q & 1);
            if (yy >= 0 && yy < p.h && xx >= 0 && xx < p.w) {
              src4[q] = x + (((long long)img * p.h + yy) * p.w + xx) * p.x_ld + c0;
              w4[q] = wq[q] * mask;
            }
          }
        }
        mbar_wait(empty_bar(s), ((kb / S) & 1) ^ 1);
        const uint32_t dst_row = smem_a + s * A_STAGE_BYTES + row_off;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (src4[q]) {
              const uint4 v = __ldg(reinterpret_cast<const uint4*>(src4[q] + 8 * j));
              const float wq = w4[q];
              acc[0] += wq * bf_lo(v.x); acc[1] += wq * bf_hi(v.x); acc[2] += wq * bf_lo(v.y); acc[3] += wq * bf_hi(v.y);
              acc[4] += wq * bf_lo(v.z); acc[5] += wq * bf_hi(v.z); acc[6] += wq * bf_lo(v.w); acc[7] += wq * bf_hi(v.w);
            }
          }
          const uint32_t d = dst_row + (((uint32_t)j ^ sw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d), "r"(pack_bf16(acc[0], acc[1])),
                       "r"(pack_bf16(acc[2], acc[3])), "r"(pack_bf16(acc[4], acc[5])), "r"(pack_bf16(acc[6], acc[7])) : "memory");
        }
        fence_proxy_async();
        mbar_arrive(full_bar(s));
      }
    }

    // ======================= epilogue (same 4 warps; warp w owns TMEM lanes 32w..32w+31) ==========
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const long long pix = (long long)oy * wo + ox;
    const float* bm_row = p.bias_map ? p.bias_map + pix * p.cout : nullptr;
#pragma unroll 1
    for (int cc = 0; cc < BN / 32; ++cc) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cc * 32), v);
      const int co0 = n0 + cc * 32;
      if (!valid || co0 >= p.cout) continue;
      const int ncol = (p.cout - co0) < 32 ? (p.cout - co0) : 32;
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float a = __uint_as_float(v[j]);
        if (j < ncol) {
          if (bm_row) a += __ldg(bm_row + co0 + j);
          a = a * __ldg(p.scale + co0 + j) + __ldg(p.shift + co0 + j);
        }
        f[j] = a;
      }
      if (p.out_dtype == PPY_BF16) {
        const bool vec = (ncol == 32) && ((p.y_ld & 7) == 0) && ((co0 & 7) == 0);
        if (p.residual) {
          const __nv_bfloat16* rr = reinterpret_cast<const __nv_bfloat16*>(p.residual) + m * p.res_ld + co0;
          if (vec && (p.res_ld & 7) == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 t = __ldg(reinterpret_cast<const uint4*>(rr) + q);
              f[8 * q + 0] += bf_lo(t.x); f[8 * q + 1] += bf_hi(t.x); f[8 * q + 2] += bf_lo(t.y); f[8 * q + 3] += bf_hi(t.y);
              f[8 * q + 4] += bf_lo(t.z); f[8 * q + 5] += bf_hi(t.z); f[8 * q + 6] += bf_lo(t.w); f[8 * q + 7] += bf_hi(t.w);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < ncol) f[j] += __bfloat162float(rr[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = apply_act(f[j], p.act);
        __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(p.y);
        const int reps = p.upsample2x ? 4 : 1;
        for (int q4 = 0; q4 < reps; ++q4) {
          long long drow = m;
          if (p.upsample2x) drow = ((long long)img * 2 * ho + 2 * oy + (q4 >> 1)) * 2 * wo + 2 * ox + (q4 & 1);
          __nv_bfloat16* dst = yb + drow * p.y_ld + co0;
          if (vec && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              reinterpret_cast<uint4*>(dst)[q] = make_uint4(pack_bf16(f[8 * q], f[8 * q + 1]), pack_bf16(f[8 * q + 2], f[8 * q + 3]),
                                                            pack_bf16(f[8 * q + 4], f[8 * q + 5]), pack_bf16(f[8 * q + 6], f[8 * q + 7]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < ncol) dst[j] = __float2bfloat16_rn(f[j]);
          }
        }
      } else {
        if (p.residual) {
          const float* rr = reinterpret_cast<const float*>(p.residual) + m * p.res_ld + co0;
#pragma unroll
          for (int j = 0; j < 32; ++j) if (j < ncol) f[j] += __ldg(rr + j);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = apply_act(f[j], p.act);
        float* yf = reinterpret_cast<float*>(p.y);
        const int reps = p.upsample2x ? 4 : 1;
        for (int q4 = 0; q4 < reps; ++q4) {
          long long drow = m;
          if (p.upsample2x) drow = ((long long)img * 2 * ho + 2 * oy + (q4 >> 1)) * 2 * wo + 2 * ox + (q4 & 1);
          float* dst = yf + drow * p.y_ld + co0;
          if (ncol == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q) reinterpret_cast<float4*>(dst)[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < ncol) dst[j] = f[j];
          }
        }
      }
    }
  } else if (warp == 4) {
    // ======================= B producer: TMA of the packed weight tile =======================
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % S;
        mbar_wait(empty_bar(s), ((kb / S) & 1) ^ 1);
        mbar_arrive_expect_tx(full_bar(s), Cfg::kBStageBytes);
        tma_load_2d(smem_b + s * Cfg::kBStageBytes, &tmap_b, full_bar(s), kb * BLOCK_K, n0);
      }
    }
  } else {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BLOCK_M, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % S;
        mbar_wait(full_bar(s), (kb / S) & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_a + s * A_STAGE_BYTES, b_addr = smem_b + s * Cfg::kBStageBytes;
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k)
          umma_bf16(tmem_base, make_smem_desc(a_addr + k * 32), make_smem_desc(b_addr + k * 32), idesc, (kb | k) ? 1u : 0u);
        umma_commit(empty_bar(s));          // frees the stage once these MMAs have read it
      }
      umma_commit(tmem_full_bar);           // accumulator complete -> epilogue
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

template <int BN, bool DCN>
int launch(const ppy_conv_params* p, int ho, int wo, cudaStream_t st) {
  using Cfg = TileCfg<BN>;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return PPY_ERR_UNSUPPORTED;
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)p->k_pad, (cuuint64_t)p->cout_pad};
  const cuuint64_t strides[1] = {(cuuint64_t)p->k_pad * 2};
  const cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)BN};
  const cuuint32_t estr[2] = {1, 1};
  CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p->weight), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { g_last_cuda_error = (int)cr; return PPY_ERR_CUDA; }
  static bool attr_done = false;
  if (!attr_done) {
    int rc = check_cuda(cudaFuncSetAttribute(conv_umma_kernel<BN, DCN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    if (rc) return rc;
    attr_done = true;
  }
  const long long M = (long long)p->n * ho * wo;
  dim3 grid((unsigned)ceil_div(p->cout, BN), (unsigned)ceil_div(M, BLOCK_M));
  conv_umma_kernel<BN, DCN><<<grid, NUM_THREADS, Cfg::kSmemBytes, st>>>(*p, ho, wo, p->k_pad / BLOCK_K, tmap);
  return check_launch();
}

template <bool DCN>
int dispatch(const ppy_conv_params* p, int ho, int wo, cudaStream_t st) {
  const int c = p->cout;
  if (c <= 32) return launch<32, DCN>(p, ho, wo, st);
  if (c <= 64) return launch<64, DCN>(p, ho, wo, st);
  if (c % 256 == 0) return launch<256, DCN>(p, ho, wo, st);
  return launch<128, DCN>(p, ho, wo, st);
}

}  // namespace
}  // namespace ppy

extern "C" {
using namespace ppy;

int ppy_conv_bf16_supported(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 && get_encode_fn() != nullptr;
}

int ppy_conv_bf16(const ppy_conv_params* p, ppy_stream_t s) {
  int ho, wo;
  int rc = validate_conv(p, 2, &ho, &wo);
  if (rc) return rc;
  PPY_REQUIRE(p->k_pad * 2 % 16 == 0);
  if (p->offset_mask) PPY_REQUIRE(p->cin % BLOCK_K == 0);
  if (!ppy_conv_bf16_supported()) return PPY_ERR_UNSUPPORTED;
  if (p->offset_mask) return dispatch<true>(p, ho, wo, as_stream(s));
  return dispatch<false>(p, ho, wo, as_stream(s));
}

}  // extern "C"
