// Fused, im2col-free DCNv2 (reference model/custom_layers.py:551-677) for whole layers on tcgen05: ONE kernel samples the NHWC
// input at the learned offsets, modulates, builds the A operand tile in shared memory and multiplies it with the 3x3 weight
// -- no [M x 9C] matrix in HBM (the reference materialises ~1.3 GB of gather temporaries per layer at bs 32).
//
//   y[M x N] = A[M x K] * W[N x K]^T,  M = n*ho*wo, N = cout (256 or 512), K = 9*C (C % 64 == 0), H*W <= 65535,
//   A[m][tap*C + c] = sigmoid(mask[m][tap]) * bilinear(x[img(m)], pos(m, tap))[c]      (zero outside the image)
//
// What bounds it, and the design that follows (B200: 148 SMs, ~6.3 KB/clk of L2 bandwidth, 128 B/clk of shared memory per SM):
//   * every (pixel, tap) needs 4 corner rows of C channels: 425 MB of L2 reads per layer at bs 32 (bf16), as much again for
//     the weights if every 128-row tile re-reads all of W.  So a CTA PAIR (cta_group::2) owns 256 rows and the FULL N: the A
//     tile is built once and multiplied into two 256-column TMEM accumulators (512 columns, M = 256 MMAs), and each CTA
//     stages only half of every weight tile -- W is read once per 256 rows.
//   * the gather must be coalesced: lane = (row of 4, 16-byte chunk of 8), so one warp load instruction covers four whole
//     128-byte corner segments (4 L1 wavefronts; one thread per row would touch 32 lines per instruction and sit on the L1
//     wavefront queue), every lane blends its four corners in fp32 registers and writes one 16-byte chunk of the swizzled A
//     tile; sixteen producer warps keep a whole K block of corner loads (64 KB) in flight per SM.
//   * sampling positions are turned ONCE per tile into a shared-memory table (per row and tap: four clamped pixel offsets and
//     four mask-multiplied bilinear weights, zero for corners outside the image), so the main loop has no transcendental,
//     no division and no branch.
//   * operands: bf16 (PAIR = false), or fp16 hi/lo pairs with three MMA groups per K block (PAIR = true, the fp32-grade
//     path; see conv_umma_impl.cuh).
//
// Warp roles (18 warps, 576 threads, <= 112 registers): warps 0-15 producers (0-7 drain the accumulators afterwards: the
// epilogue applies scale/shift, residual, activation and stores bf16 / pair / fp32 rows), warp 16 weight TMA, warp 17 TMEM
// allocation + MMA issue (leader CTA).  One tile per CTA pair: M / 256 pairs (46 of 74 at bs 32 x 19x19).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"
#include "umma_ptx.cuh"

namespace ppy {

int validate_conv(const ppy_conv_params* p, int elem_bytes, int* ho, int* wo);

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int DB_M = 128;                         // rows per CTA (256 per pair)
constexpr int DB_K = 64;                          // 16-bit elements per K block = one 128-byte swizzle row
constexpr int D_PROD_WARPS = 16, D_TMA_WARP = 16, D_MMA_WARP = 17, D_THREADS = 18 * 32;
constexpr int D_EPI_WARPS = 8;
constexpr int D_A_TILE = DB_M * DB_K * 2;         // 16 KB
constexpr int D_B_HALF_ROWS = 128;                // rows of a 256-wide weight tile staged by one CTA
constexpr int D_B_TILE = D_B_HALF_ROWS * DB_K * 2;   // 16 KB: this CTA's half of one 256-column weight tile
constexpr int D_MAX_NT = 2;                       // 256-column accumulators (N <= 512)
constexpr int D_TAPS = 9;
constexpr int D_ENTRY_BYTES = 24;                 // per (row, tap): 4 x u16 pixel offsets (in-image) | 4 x f32 weights
constexpr int D_TABLE_BYTES = DB_M * D_TAPS * D_ENTRY_BYTES;

template <bool PAIR> struct DcnCfg {
  static constexpr int kPlanes = PAIR ? 2 : 1;
  static constexpr int kAStage = kPlanes * D_A_TILE;
  static constexpr int kBStage = kPlanes * D_MAX_NT * D_B_TILE;
  static constexpr int kStages = PAIR ? 2 : 4;
  static constexpr int kSmem = kStages * (kAStage + kBStage) + D_TABLE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static_assert(kSmem <= 232448, "shared memory budget");
};

__device__ __forceinline__ uint32_t d_pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float d_bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float d_bf_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ float2 d_h2f(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }
__device__ __forceinline__ void d_split(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__host__ __device__ constexpr uint32_t d_idesc(int m, int n, bool f16) {
  return (1u << 4) | (f16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mbar_arrive_any(uint32_t local_bar, uint32_t cluster_bar_rank0, bool leader) {
  if (leader) mbar_arrive(local_bar); else mbar_arrive_cluster(cluster_bar_rank0);
}

template <bool PAIR>
__global__ void __launch_bounds__(D_THREADS, 1)
dcn_umma_kernel(const ppy_conv_params p, const int ho, const int wo, const int num_kb, const int nt_count,
                const __grid_constant__ CUtensorMap tmap_b) {
  using Cfg = DcnCfg<PAIR>;
  constexpr int S = Cfg::kStages, PL = Cfg::kPlanes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t smem_a = smem_base, smem_b = smem_base + S * Cfg::kAStage;
  const uint32_t tab_off = S * (Cfg::kAStage + Cfg::kBStage);
  uint8_t* table = gen_base + tab_off;
  const uint32_t bars = smem_base + tab_off + D_TABLE_BYTES;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen_base + tab_off + D_TABLE_BYTES + (2 * S + 1) * 8);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
  const uint32_t tmem_full = bars + 8u * (2 * S);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int cta_rank = (int)(blockIdx.x & 1);
  const bool leader = cta_rank == 0;
  const long long M = (long long)p.n * ho * wo;
  const long long m_base = ((long long)(blockIdx.x >> 1) * 2 + cta_rank) * DB_M;      // first row of this CTA
  const int hw_out = ho * wo;
  const int tmem_cols = nt_count == 1 ? 256 : 512;

  if (warp == D_TMA_WARP && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
  if (tid == 0) {
    // full: every producer warp of BOTH CTAs arrives on the leader's barrier + the leader's TMA thread (expect_tx)
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 2 * D_PROD_WARPS + 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == D_MMA_WARP) tmem_alloc2(smem_u32(const_cast<uint32_t*>(tmem_slot)), (uint32_t)tmem_cols);

  // ---- sampling table of this CTA's 128 rows (all threads): per (row, tap) four clamped in-image pixel offsets + four weights
  for (int e = tid; e < DB_M * D_TAPS; e += D_THREADS) {
    const int r = e / D_TAPS, tap = e % D_TAPS;
    const long long m = m_base + r;
    uint16_t off[4] = {0, 0, 0, 0};
    float w4[4] = {0.f, 0.f, 0.f, 0.f};
    if (m < M) {
      const int pix = (int)(m % hw_out);
      const int oy = pix / wo, ox = pix % wo;
      const float* om = p.offset_mask + m * p.om_ld;
      const float dy = __ldg(om + 2 * tap), dx = __ldg(om + 2 * tap + 1), ml = __ldg(om + 2 * D_TAPS + tap);
      const float mask = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-ml)));
      const float py = (float)(oy * p.stride - p.pad + tap / 3) + dy, px = (float)(ox * p.stride - p.pad + tap % 3) + dx;
      const float fy = floorf(py), fx = floorf(px);
      const float ly = py - fy, lx = px - fx, hy = 1.f - ly, hx = 1.f - lx;
      const int y0 = (int)fminf(fmaxf(fy, -2.f), (float)p.h), x0 = (int)fminf(fmaxf(fx, -2.f), (float)p.w);
      const float wq[4] = {hy * hx, hy * lx, ly * hx, ly * lx};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int yy = y0 + (q >> 1), xx = x0 + (q & 1);
        const bool ok = yy >= 0 && yy < p.h && xx >= 0 && xx < p.w;
        w4[q] = ok ? wq[q] * mask : 0.f;
        off[q] = (uint16_t)(min(max(yy, 0), p.h - 1) * p.w + min(max(xx, 0), p.w - 1));
      }
    }
    uint32_t* dst = reinterpret_cast<uint32_t*>(table + e * D_ENTRY_BYTES);
    dst[0] = (uint32_t)off[0] | ((uint32_t)off[1] << 16);
    dst[1] = (uint32_t)off[2] | ((uint32_t)off[3] << 16);
    dst[2] = __float_as_uint(w4[0]); dst[3] = __float_as_uint(w4[1]); dst[4] = __float_as_uint(w4[2]); dst[5] = __float_as_uint(w4[3]);
  }
  tc_fence_before();
  cluster_sync_all();                            // barriers of both CTAs initialised, tables written, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int cb_per_tap = p.cin / DB_K;

  if (warp < D_PROD_WARPS) {
    // =====================================================================================
    // producers: warp w owns rows 8w .. 8w+7 of the tile, two passes of four rows; lane = (row of the pass, 16-byte chunk).
    // All corner loads of a K block (8 x 16 bytes per lane; pair operands: hi and lo planes, 16) are issued before the stage's
    // empty barrier is even waited for.  (Measured alternative: two groups of eight warps building alternate K blocks with
    // twice the loads in flight per warp -- 10 % slower; the kernel is bound by what one SM can ingest from L2 per K block,
    // 64 KB of corners + 32 KB of weights, not by load latency.)
    // =====================================================================================
    const int rsub = lane >> 3, j = lane & 7;
    const uint16_t* x16 = reinterpret_cast<const uint16_t*>(p.x);
    const uint32_t full_rank0 = map_to_cta(full_bar(0), 0);
    int row[2];
    long long img_pix[2];                        // first pixel of the row's image
    row[0] = warp * 8 + rsub; row[1] = row[0] + 4;
#pragma unroll
    for (int ps = 0; ps < 2; ++ps) {
      const long long m = m_base + row[ps];
      img_pix[ps] = (m < M ? m / hw_out : 0) * (long long)p.h * p.w;
    }
    uint32_t offp[2][2];
    float wq[2][4];
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % S;
      const int tap = kb / cb_per_tap, c0 = (kb % cb_per_tap) * DB_K;
      if (c0 == 0) {                              // new tap: this lane's two table entries
#pragma unroll
        for (int ps = 0; ps < 2; ++ps) {
          const uint32_t* e = reinterpret_cast<const uint32_t*>(table + (row[ps] * D_TAPS + tap) * D_ENTRY_BYTES);
          offp[ps][0] = e[0]; offp[ps][1] = e[1];
          wq[ps][0] = __uint_as_float(e[2]); wq[ps][1] = __uint_as_float(e[3]); wq[ps][2] = __uint_as_float(e[4]); wq[ps][3] = __uint_as_float(e[5]);
        }
      }
      uint4 v[2][4], v2[PAIR ? 2 : 1][PAIR ? 4 : 1];
#pragma unroll
      for (int ps = 0; ps < 2; ++ps) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t o = (offp[ps][q >> 1] >> ((q & 1) * 16)) & 0xFFFFu;
          const uint16_t* src = x16 + (img_pix[ps] + o) * p.x_ld + c0 + j * 8;
          v[ps][q] = __ldg(reinterpret_cast<const uint4*>(src));
          if (PAIR) v2[PAIR ? ps : 0][PAIR ? q : 0] = __ldg(reinterpret_cast<const uint4*>(src + p.x_plane));
        }
      }
      mbar_wait(empty_bar(s), ((kb / S) & 1) ^ 1);
      const uint32_t a_stage = smem_a + s * Cfg::kAStage;
#pragma unroll
      for (int ps = 0; ps < 2; ++ps) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float wgt = wq[ps][q];
          const uint4 t = v[ps][q];
          if (PAIR) {
            const uint4 u = v2[PAIR ? ps : 0][PAIR ? q : 0];
            float2 a, b;
            a = d_h2f(t.x); b = d_h2f(u.x); acc[0] += wgt * (a.x + b.x); acc[1] += wgt * (a.y + b.y);
            a = d_h2f(t.y); b = d_h2f(u.y); acc[2] += wgt * (a.x + b.x); acc[3] += wgt * (a.y + b.y);
            a = d_h2f(t.z); b = d_h2f(u.z); acc[4] += wgt * (a.x + b.x); acc[5] += wgt * (a.y + b.y);
            a = d_h2f(t.w); b = d_h2f(u.w); acc[6] += wgt * (a.x + b.x); acc[7] += wgt * (a.y + b.y);
          } else {
            acc[0] += wgt * d_bf_lo(t.x); acc[1] += wgt * d_bf_hi(t.x); acc[2] += wgt * d_bf_lo(t.y); acc[3] += wgt * d_bf_hi(t.y);
            acc[4] += wgt * d_bf_lo(t.z); acc[5] += wgt * d_bf_hi(t.z); acc[6] += wgt * d_bf_lo(t.w); acc[7] += wgt * d_bf_hi(t.w);
          }
        }
        const uint32_t d = a_stage + (uint32_t)row[ps] * 128u + (((uint32_t)j ^ (uint32_t)(row[ps] & 7)) << 4);
        if (PAIR) {
          uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
          d_split(acc[0], acc[1], h0, l0); d_split(acc[2], acc[3], h1, l1); d_split(acc[4], acc[5], h2, l2); d_split(acc[6], acc[7], h3, l3);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d), "r"(h0), "r"(h1), "r"(h2), "r"(h3) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d + (uint32_t)D_A_TILE), "r"(l0), "r"(l1), "r"(l2), "r"(l3) : "memory");
        } else {
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d), "r"(d_pack_bf16(acc[0], acc[1])), "r"(d_pack_bf16(acc[2], acc[3])),
                       "r"(d_pack_bf16(acc[4], acc[5])), "r"(d_pack_bf16(acc[6], acc[7])) : "memory");
        }
      }
      fence_proxy_async();                       // generic-proxy writes of the A tile -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive_any(full_bar(s), full_rank0 + 8u * s, leader);
    }
  } else if (warp == D_TMA_WARP) {
    // =====================================================================================
    // weight TMA: per K block and 256-column tile this CTA's 128-row half (two 64-row boxes; pair operands: hi then lo rows)
    // =====================================================================================
    const uint32_t tx_bytes = (uint32_t)(2 * PL * nt_count * D_B_TILE);          // both CTAs' bytes land on the leader's barrier
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % S;
      mbar_wait(empty_bar(s), ((kb / S) & 1) ^ 1);
      if (elect_one()) {
        const uint32_t bar = map_to_cta(full_bar(s), 0);
        if (leader) mbar_arrive_expect_tx(full_bar(s), tx_bytes);
        for (int pl = 0; pl < PL; ++pl)
          for (int nt = 0; nt < nt_count; ++nt)
#pragma unroll
            for (int hb = 0; hb < 2; ++hb)
              tma2_load_2d(smem_b + s * Cfg::kBStage + (pl * D_MAX_NT + nt) * D_B_TILE + hb * (D_B_TILE / 2), &tmap_b, bar, kb * DB_K,
                           pl * p.cout_pad + nt * 256 + cta_rank * D_B_HALF_ROWS + hb * 64);
      }
      __syncwarp();
    }
  } else if (warp == D_MMA_WARP) {
    // =====================================================================================
    // MMA issuer (leader CTA): per K block, per 256-column accumulator, 4 (pair operands: 12) M=256 MMAs
    // =====================================================================================
    if (leader) {
      constexpr uint32_t idesc = d_idesc(2 * DB_M, 256, PAIR);
      constexpr int COMBOS = PAIR ? 3 : 1;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % S;
        mbar_wait(full_bar(s), (kb / S) & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_a + s * Cfg::kAStage, b_addr = smem_b + s * Cfg::kBStage;
        if (elect_one()) {
          for (int nt = 0; nt < nt_count; ++nt) {
#pragma unroll
            for (int cb = 0; cb < COMBOS; ++cb) {
              const uint32_t a_pl = a_addr + (cb == 2 ? D_A_TILE : 0), b_pl = b_addr + ((cb == 1 ? D_MAX_NT : 0) + nt) * D_B_TILE;
#pragma unroll
              for (int k = 0; k < DB_K / 16; ++k)
                umma2_bf16(tmem_base + (uint32_t)(nt * 256), make_smem_desc(a_pl + k * 32), make_smem_desc(b_pl + k * 32), idesc, (kb | cb | k) ? 1u : 0u);
            }
          }
          umma_commit2(empty_bar(s));
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit2(tmem_full);
      __syncwarp();
    }
  }

  if (warp < D_EPI_WARPS) {
    // =====================================================================================
    // epilogue (warps 0-7 after their producer loop): lane = tile row; warp w reads TMEM lanes 32(w&3).. and the 32-column
    // groups (w>>2), (w>>2)+2, ..; scale/shift, residual, activation; each lane stores 32 channels of its row
    // =====================================================================================
    const int quarter = warp & 3, half = warp >> 2;
    const int r = quarter * 32 + lane;
    const long long m = m_base + r;
    const float slope = p.act == PPY_ACT_RELU ? 0.f : (p.act == PPY_ACT_LEAKY ? 0.1f : 1.f);
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int groups = (p.cout + 31) / 32;
    for (int g = half; g < groups; g += 2) {
      uint32_t v[32];
      tmem_ld32_nowait(t_row + (uint32_t)(g * 32), v);
      tmem_wait_ld();
      if (m >= M) continue;
      const int co0 = g * 32;
#pragma unroll
      for (int q = 0; q < 4; ++q) {                 // 8 channels at a time
        const int co = co0 + q * 8;
        if (co >= p.cout) break;
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = co + e;
          f[e] = c < p.cout ? __uint_as_float(v[q * 8 + e]) * __ldg(p.scale + c) + __ldg(p.shift + c) : 0.f;
        }
        const bool full8 = co + 8 <= p.cout;
        if (p.residual) {
          if (p.out_dtype == PPY_F16X2) {
            const __half* rp = reinterpret_cast<const __half*>(p.residual) + (size_t)m * p.res_ld + co;
            for (int e = 0; e < 8 && co + e < p.cout; ++e) f[e] += __half2float(rp[e]) + __half2float(rp[p.res_plane + e]);
          } else if (p.out_dtype == PPY_BF16) {
            const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + (size_t)m * p.res_ld + co;
            for (int e = 0; e < 8 && co + e < p.cout; ++e) f[e] += __bfloat162float(rp[e]);
          } else {
            const float* rp = reinterpret_cast<const float*>(p.residual) + (size_t)m * p.res_ld + co;
            for (int e = 0; e < 8 && co + e < p.cout; ++e) f[e] += rp[e];
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], f[e] * slope);
        if (p.out_dtype == PPY_F16X2) {
          __half* d = reinterpret_cast<__half*>(p.y) + (size_t)m * p.y_ld + co;
          float mx = 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) mx = fmaxf(mx, fabsf(f[e]));
          if (!(mx <= 65504.f) && p.overflow) *p.overflow = 1;
          if (full8) {
            uint4 hi, lo;
            d_split(f[0], f[1], hi.x, lo.x); d_split(f[2], f[3], hi.y, lo.y); d_split(f[4], f[5], hi.z, lo.z); d_split(f[6], f[7], hi.w, lo.w);
            *reinterpret_cast<uint4*>(d) = hi;
            *reinterpret_cast<uint4*>(d + p.y_plane) = lo;
          } else {
            for (int e = 0; co + e < p.cout; ++e) { const __half h = __float2half_rn(f[e]); d[e] = h; d[p.y_plane + e] = __float2half_rn(f[e] - __half2float(h)); }
          }
        } else if (p.out_dtype == PPY_BF16) {
          __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(p.y) + (size_t)m * p.y_ld + co;
          if (full8) *reinterpret_cast<uint4*>(d) = make_uint4(d_pack_bf16(f[0], f[1]), d_pack_bf16(f[2], f[3]), d_pack_bf16(f[4], f[5]), d_pack_bf16(f[6], f[7]));
          else for (int e = 0; co + e < p.cout; ++e) d[e] = __float2bfloat16_rn(f[e]);
        } else {
          float* d = reinterpret_cast<float*>(p.y) + (size_t)m * p.y_ld + co;
          if (full8) { reinterpret_cast<float4*>(d)[0] = make_float4(f[0], f[1], f[2], f[3]); reinterpret_cast<float4*>(d)[1] = make_float4(f[4], f[5], f[6], f[7]); }
          else for (int e = 0; co + e < p.cout; ++e) d[e] = f[e];
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();                             // nobody leaves while the peer may still read its shared memory / TMEM
  if (warp == D_MMA_WARP) tmem_dealloc2(tmem_base, (uint32_t)tmem_cols);
}

template <bool PAIR>
int dcn_launch(const ppy_conv_params* p, int ho, int wo, cudaStream_t st) {
  using Cfg = DcnCfg<PAIR>;
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return PPY_ERR_UNSUPPORTED;
    enc = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  CUtensorMap tmap_b;
  const cuuint64_t dims[2] = {(cuuint64_t)p->k_pad, (cuuint64_t)(PAIR ? 2 : 1) * p->cout_pad};
  const cuuint64_t strides[1] = {(cuuint64_t)p->k_pad * 2};
  const cuuint32_t box[2] = {(cuuint32_t)DB_K, 64};
  const cuuint32_t estr[2] = {1, 1};
  CUresult cr = enc(&tmap_b, PAIR ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p->weight), dims,
                    strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { g_last_cuda_error = (int)cr; return PPY_ERR_CUDA; }
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    int rc = check_cuda(cudaFuncSetAttribute(dcn_umma_kernel<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    if (rc) return rc;
  }
  const long long M = (long long)p->n * ho * wo;
  const long long pairs = ceil_div(M, 2 * DB_M);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  cfg.blockDim = dim3(D_THREADS);
  cfg.dynamicSmemBytes = Cfg::kSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int rc = check_cuda(cudaLaunchKernelEx(&cfg, dcn_umma_kernel<PAIR>, *p, ho, wo, p->k_pad / DB_K, p->cout / 256, tmap_b));
  if (rc) return rc;
  return check_launch();
}

}  // namespace

// The whole-layer fused DCNv2 kernel when the layer fits it (3x3, C % 64 == 0, cout 256 or 512, H*W <= 65535, plain epilogue);
// returns 1 ("not applicable") otherwise so the caller can fall back to the generic producer-mode kernel.
int dcn_umma_try(const ppy_conv_params* p, int ho, int wo, bool pair, cudaStream_t st) {
  static const bool off = getenv("PPY_NO_DCN2") != nullptr;
  if (off || !p->offset_mask) return 1;
  if (p->kh != 3 || p->kw != 3 || p->cin % DB_K || p->k_pad != 9 * p->cin) return 1;
  if (p->cout != 256 && p->cout != 512) return 1;
  if ((long long)p->h * p->w > 65535 || (long long)p->n * p->h * p->w * p->x_ld >= 0x7FFFFFFFll) return 1;
  if (p->bias_map || p->coord_w || p->upsample2x || p->accumulate) return 1;
  const int esz = p->out_dtype == PPY_F32 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(p->y) & 15) || (p->y_ld * esz) % 16 || (reinterpret_cast<uintptr_t>(p->x) & 15) || (p->x_ld * 2) % 16) return 1;
  if (pair && ((p->x_plane * 2) % 16 || (p->out_dtype == PPY_F16X2 && (p->y_plane * 2) % 16))) return 1;
  if (p->act == PPY_ACT_MISH) return 1;
  return pair ? dcn_launch<true>(p, ho, wo, st) : dcn_launch<false>(p, ho, wo, st);
}

}  // namespace ppy
