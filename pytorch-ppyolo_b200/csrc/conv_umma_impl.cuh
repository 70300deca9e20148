// bf16 implicit-GEMM convolution on the sm_100a tensor cores (tcgen05.mma, fp32 accumulators in TMEM),
// with fused scale/shift (+CoordConv bias map or rank-2 term, +residual), activation and optional 2x nearest upsample
// in the epilogue.  Also the fused DCNv2 kernel: the same GEMM pipeline with a producer that bilinearly
// samples the NHWC input at the learned offsets and writes the modulated bf16 A tile straight into the
// swizzled shared-memory stage -- no im2col / gather temporaries in HBM (reference
// model/custom_layers.py:571-676 materialises ~1.3 GB of them per layer at bs=32).
//
//   GEMM view: D[M x N] = A[M x K] * B[N x K]^T,  M = n*ho*wo pixels, N = cout, K = kh*kw*cin,
//   K index = (ky*kw + kx)*cin + c.  CTA tile 128 x BLOCK_N x 64.
//
// PERSISTENT, warp-specialised: one CTA per SM (or one CTA pair per TPC, template flag CTA2: cta_group::2, 256 x BLOCK_N
// tiles, B split across the pair) walks tiles (n fastest) with stride gridDim.x; the TMEM accumulator is double buffered so the
// epilogue of tile i overlaps the main loop of tile i+1.  Launched with programmatic stream serialization: the prologue
// overlaps the previous kernel's tail, griddepcontrol.wait precedes the first global access.
//
//   warps 0-7  epilogue.  EPI_SLAB: warp w owns TMEM lanes 32(w&3)..+31 and every other 32-column sub-tile; tcgen05.ld (next
//              sub-tile prefetched) -> XOR-swizzled warp-private smem slab -> coalesced pass (lane = 8 channels of a row):
//              CoordConv bias map, scale/shift, residual (16-byte loads one sub-tile ahead), activation, 16-byte stores.
//              EPI_TMA (bf16 layers with short K, the HBM-bound ones): every warp works alone on [32 rows x 64 channels] boxes:
//              its lane 0 fetches the residual box by TMA two boxes ahead, lane = tile row does the math straight from the TMEM
//              registers (CoordConv rank-2 term, scale/shift from a per-warp smem table, residual, activation) into a swizzled
//              output box that leaves by TMA store -- no global-memory instruction, no CTA-wide barrier.
//   warps 8-11 A producers, present only in MODE gather (cin % 64 != 0 leftovers: eight 16-byte cp.async per thread and K block,
//              zero-fill outside the image, straight into the swizzled stage) and MODE dcn (thread = tile row, bilinear
//              sample x mask in fp32 -> bf16 st.shared); idle in the TMA modes.
//   next warp  TMA producer (warp-uniform loop, elect.sync lane issues): weight tile(s) per stage, plus the A operand:
//              tma_a      1x1 stride-1: plain [128 x 64] box of the NHWC matrix
//              tma_patch  3x3 stride-1: 4-D box {64 ch, 16, 8, 1} per (tap, channel block) at pixel offset (kx-1, ky-1), zero halo
//                         from TMA's out-of-bounds fill, the 128 tile rows are a 16x8 pixel patch
//              tma_slab   3x3 stride-1, cout <= 128: 8x18-pixel slab per (channel block, kx) + the three weight tiles of its
//                         taps; tap ky reads the same slab at +ky*1024 bytes (see MODE_TMA_SLAB below)
//              tma_im2col any other k x k / stride: im2col-mode tensor map, 128 consecutive output pixels per box, filter
//                         offset in the instruction, padding zero-filled by the copy engine
//   last warp  TMEM allocator + MMA issuer (warp-uniform loop): 4 x tcgen05.mma (K=16) per 64-wide K block (12 per slab
//              stage), tcgen05.commit to the stage's empty barrier, one commit per tile to tmem_full
//
// Shared-memory operand layout is the canonical K-major SWIZZLE_128B one: row r of a stage lives at
// r*128 bytes, its 16-byte chunk j at ((j ^ (r & 7)) << 4); 8-row groups are 1024 bytes apart (SBO).
//
// This file is compiled twice: PPY_UMMA_SPLIT == 0 -> the bf16 kernels (ppy_conv_bf16), PPY_UMMA_SPLIT == 1 -> the fp32-grade
// kernels (ppy_conv_f16x2): every operand is an fp16 hi/lo PAIR (two planes, v = hi + lo to 2^-22), a stage holds the hi and
// the lo tile of A and of B, and each K block issues three MMA groups -- hi*hi, hi*lo, lo*hi -- into the same fp32 accumulator
// (the dropped lo*lo term is 2^-22 relative): fp32-FMA-chain accuracy at a third of the bf16 tensor rate.  The epilogue
// re-splits its fp32 results into pairs.  Only the slab epilogue exists in split mode.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"
#include "umma_ptx.cuh"

#ifndef PPY_UMMA_SPLIT
#define PPY_UMMA_SPLIT 0
#endif
// PPY_UMMA_K32 == 1 (third compilation, split mode only: ppy_conv_f16x2_k32): K blocks of 32 elements -- 64-byte operand rows in
// the SWIZZLE_64B layout -- for the 3x3 convs with 32 input channels (stem conv1_2 / conv1_3): one K block is exactly one tap, so
// no MAC multiplies a structural zero (the pixel-pair reinterpretation that feeds those layers to the 64-wide kernel wastes half).
// Only the slab loader is dispatched in that build.
#ifndef PPY_UMMA_K32
#define PPY_UMMA_K32 0
#endif

namespace ppy {

int validate_conv(const ppy_conv_params* p, int elem_bytes, int* ho, int* wo);
int dcn_umma_try(const ppy_conv_params* p, int ho, int wo, bool pair, cudaStream_t st);     // dcn_umma.cu; 1 = not applicable

namespace {

constexpr bool SPLIT = PPY_UMMA_SPLIT != 0;       // fp16 hi/lo pair operands, three MMA groups per K block
constexpr int PLANES = SPLIT ? 2 : 1;
constexpr int BLOCK_M = 128;
constexpr bool K32 = PPY_UMMA_K32 != 0;
constexpr int BLOCK_K = K32 ? 32 : 64;            // 16-bit elements of one operand row = one swizzle row (128 bytes; K32: 64 bytes)
constexpr int ROW_BYTES = BLOCK_K * 2;
constexpr int ATOM_BYTES = 8 * ROW_BYTES;         // swizzle atom: 8 rows (SBO of the shared-memory descriptors)
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int EPI_WARPS = 8;                      // two warps per TMEM lane quarter, alternating sub-tiles
constexpr int PRODUCER_WARP0 = EPI_WARPS;
constexpr int EPI_SLAB = 0, EPI_TMA = 1;          // epilogue variants (see the kernel header)
constexpr int GROUP_COLS = 64;                    // EPI_TMA: residual / output move as [32 rows x 64 ch] bf16 boxes (4 KB), one warp each
constexpr int BOX_ROWS = 32;
constexpr int BOX_BYTES = BOX_ROWS * GROUP_COLS * 2;
constexpr int EPI_TABLE_FLOATS = 4 * GROUP_COLS;  // per epilogue warp: scale | shift | wx | wy of its current 64 columns
constexpr int MODE_GATHER = 0, MODE_TMA_A = 1, MODE_DCN = 2, MODE_TMA_PATCH = 3, MODE_TMA_IM2COL = 4, MODE_TMA_SLAB = 5;
__host__ __device__ constexpr bool mode_is_tma(int mode) {
  return mode == MODE_TMA_A || mode == MODE_TMA_PATCH || mode == MODE_TMA_IM2COL || mode == MODE_TMA_SLAB;
}
// 3x3 stride-1 convs: the 128 tile rows are a pixel patch of one image -- 16 wide x 8 tall (MODE_TMA_PATCH, one box per tap), or
// 8 wide x 16 tall (MODE_TMA_SLAB).  SLAB: one pipeline stage holds, for one 64-channel block and one kx, the 8 x 18 pixel slab
// (two halo rows) plus the three weight tiles of taps (ky = 0..2, kx); an 8-pixel slab row is exactly one 1024-byte swizzle
// atom, so the A operand of tap ky is the SAME slab read at byte offset ky * 1024 -- each input pixel enters shared memory
// 3.4 times per tile instead of 9, and there is one barrier round trip per THREE taps.
// Warp roles.  TMA-fed modes: 8 epilogue warps + TMA producer + MMA issuer = 10 warps (registers are allocated per group of four
// warps, so 9..12 warps may use 168 registers per thread where 13..16 are capped at 128: the epilogues want them).  gather / dcn
// modes insert four A-producer warps after the epilogue warps (14 warps).
__host__ __device__ constexpr int producer_warps(int mode) { return mode_is_tma(mode) ? 0 : 4; }
__host__ __device__ constexpr int tma_warp(int mode) { return EPI_WARPS + producer_warps(mode); }
__host__ __device__ constexpr int mma_warp(int mode) { return tma_warp(mode) + 1; }
__host__ __device__ constexpr int num_threads(int mode) { return 32 * (mma_warp(mode) + 1); }
__host__ __device__ constexpr bool mode_is_patchy(int mode) { return mode == MODE_TMA_PATCH || mode == MODE_TMA_SLAB; }
__host__ __device__ constexpr int patch_w(int mode) { return mode == MODE_TMA_SLAB ? 8 : 16; }
__host__ __device__ constexpr int patch_h(int mode) { return mode == MODE_TMA_SLAB ? 16 : 8; }
constexpr int SLAB_ROWS = 18;                     // patch_h + 2 halo rows
constexpr int SLAB_BYTES = SLAB_ROWS * 8 * ROW_BYTES;   // 18 rows x 8 pixels x one operand row
constexpr int SUB = 32;                           // epilogue sub-tile columns (= one tcgen05.ld.x32)
constexpr int ST_LD = SUB;                        // floats per staged row; 16-byte chunks XOR-swizzled by (row & 7)
constexpr int STAGING_BYTES = EPI_WARPS * 32 * ST_LD * 4;
// Split mode, K chunks.  The tensor core adds every MMA (K = 16) into the fp32 accumulator with truncation, a bias that grows
// linearly with the number of MMAs per accumulator (measured 1.5e-5 of the output scale at K = 4608 with one accumulator).  So a
// TMEM accumulator only ever holds the partial sum of one K CHUNK (chunk_kb pipeline iterations, <= 72 MMAs); the epilogue warps
// drain every chunk into REGISTER accumulators with round-to-nearest adds while the MMA issuer fills the other TMEM buffer.
// BLOCK_N <= 128 in split mode (64 accumulator registers per epilogue thread).
// Tiles whose whole K fits one chunk (K <= 256 of a 1x1 conv: the HBM-bound layers) skip the register stage (template flag
// CHUNKED off) and keep BLOCK_N up to 256.
__host__ __device__ constexpr int chunk_kb(int mode) { return mode == MODE_TMA_SLAB ? 2 : 4; }

template <int BN, int MODE, int EPI, bool CTA2 = false> struct TileCfg {
  // CTA2 (cta_group::2 pair, 256 x BN tile): each CTA stages its own 128 rows of A and HALF of the B tile, so stages are smaller
  // and the ring deeper.  EPI_TMA is only dispatched for small K (<= 512 at BLOCK_N 256, <= 1152 below), so fewer stages suffice
  // there and free shared memory for the tile buffers.  SLAB stages hold a pixel slab and three weight tiles.
  static constexpr bool kSlab = MODE == MODE_TMA_SLAB;
  static constexpr int kBTileBytes = BN * BLOCK_K * 2 / (CTA2 ? 2 : 1);            // one [BN x 64] weight tile (this CTA's half)
  static constexpr int kATileBytes = kSlab ? SLAB_BYTES : A_STAGE_BYTES;            // one plane of the A operand of a stage
  static constexpr int kBPlaneBytes = kSlab ? 3 * kBTileBytes : kBTileBytes;        // one plane of the B operand of a stage
  static constexpr int kAStageBytes = PLANES * kATileBytes;                         // split mode: hi tile | lo tile
  static constexpr int kBStageBytes = PLANES * kBPlaneBytes;
  // TMA epilogue, per warp: residual boxes (double buffered; single in SLAB mode, whose 3x3 layers rarely carry a residual and
  // whose stages need the room) + one output box + the per-channel table
  static constexpr int kResBufs = kSlab ? 1 : 2;
  static constexpr int kEpiWarpBytes = (kResBufs + 1) * BOX_BYTES;
  static constexpr int kEpiBytes = EPI == EPI_TMA ? EPI_WARPS * (kEpiWarpBytes + EPI_TABLE_FLOATS * 4) : STAGING_BYTES;
  static constexpr int kParamBytes = 320;      // shared-memory copy of the parameter block for the out-of-line generic epilogue
  static constexpr int kFixedBytes = kEpiBytes + 1024 /*align slack*/ + 512 /*barriers*/ + kParamBytes;
  static constexpr int kFit = (232448 - kFixedBytes) / (kAStageBytes + kBStageBytes);
  static_assert(kFit == (232448 - (kFixedBytes - kParamBytes)) / (kAStageBytes + kBStageBytes), "the parameter copy must not cost a pipeline stage");
  static constexpr int kWant = CTA2 ? (EPI == EPI_TMA ? (BN == 256 ? 4 : 6) : (BN == 256 ? 6 : 8))
                                    : (EPI == EPI_TMA ? (BN == 256 ? 3 : (BN == 128 ? 4 : 6)) : (BN == 256 ? 4 : (BN == 128 ? 6 : 8)));
  static constexpr int kStages = (kSlab || SPLIT) ? (kFit < 4 ? kFit : 4) : (kFit < kWant ? kFit : kWant);
  static constexpr int kCpLag = kStages - 2;               // cp.async groups in flight per producer thread
  static constexpr int kTmemCols = 2 * BN;                 // double-buffered accumulator; power of two >= 64
  // EPI_SLAB: 8 warp-private fp32 slabs.  EPI_TMA: per warp 2 residual boxes + 1 output box + its per-channel table
  static constexpr int kSmemBytes = kStages * (kAStageBytes + kBStageBytes) + kFixedBytes;
  static_assert(kStages >= 2 && kSmemBytes <= 232448, "shared memory budget");
};

// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) at [4,6), a/b format BF16 (1) at
// [7,10)/[10,13), K-major A and B, N>>3 at [17,23), M>>4 at [24,29).
// (split mode: a/b format F16 = 0)
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (SPLIT ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// shared-memory operand descriptor of this build's swizzle (K-major; 128-byte rows / SWIZZLE_128B, or 64-byte rows / SWIZZLE_64B)
__device__ __forceinline__ uint64_t mk_desc(uint32_t smem_addr) { return K32 ? make_smem_desc_sw64(smem_addr) : make_smem_desc(smem_addr); }

__device__ __forceinline__ float bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// fp16 pair helpers (split mode): v = hi + lo; two values per 32-bit word like the bf16 packing
__device__ __forceinline__ float2 h2f(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }
__device__ __forceinline__ void split_pack(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// 8 consecutive channels -> one 16-byte vector of the hi plane and one of the lo plane; flags values beyond the fp16 range
__device__ __forceinline__ void split8(const float4& a, const float4& b, uint4& hi, uint4& lo, int* overflow) {
  split_pack(a.x, a.y, hi.x, lo.x); split_pack(a.z, a.w, hi.y, lo.y);
  split_pack(b.x, b.y, hi.z, lo.z); split_pack(b.z, b.w, hi.w, lo.w);
  const float mx = fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))),
                         fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))));
  if (!(mx <= 65504.f) && overflow) *overflow = 1;      // also catches NaN
}
__device__ __forceinline__ void add_pair8(float4& a, float4& b, const uint4& hi, const uint4& lo) {
  float2 t, u;                                          // hi + lo is exact in fp32: one rounding per add, like an fp32 residual
  t = h2f(hi.x); u = h2f(lo.x); a.x += t.x + u.x; a.y += t.y + u.y;
  t = h2f(hi.y); u = h2f(lo.y); a.z += t.x + u.x; a.w += t.y + u.y;
  t = h2f(hi.z); u = h2f(lo.z); b.x += t.x + u.x; b.y += t.y + u.y;
  t = h2f(hi.w); u = h2f(lo.w); b.z += t.x + u.x; b.w += t.y + u.y;
}

// nearest x2 upsample fused into the store: one output pixel -> its 2x2 block (aligned 8-channel vectors)
// (arguments by value: a reference to the parameter block would make every access a local-memory load -- an L2 round trip next
// to 200+ KB of shared memory, see epilogue_acc)
__device__ __noinline__ void store_upsampled(void* y, int y_ld, long long y_plane, int out_dtype, int* overflow, int m, int co, int ho, int wo,
                                            float4 a, float4 b) {
  const unsigned hw_out = (unsigned)(ho * wo);
  const unsigned pix = (unsigned)m % hw_out, img = (unsigned)m / hw_out;
  const unsigned oy = pix / (unsigned)wo, ox = pix % (unsigned)wo;
  const size_t r0 = ((size_t)img * 2 * ho + 2 * oy) * 2 * wo + 2 * ox;
  const size_t rows[4] = {r0, r0 + 1, r0 + 2 * (size_t)wo, r0 + 2 * (size_t)wo + 1};
  if (out_dtype == PPY_F16X2) {
    uint4 hi, lo;
    split8(a, b, hi, lo, overflow);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      __half* d = reinterpret_cast<__half*>(y) + rows[q] * y_ld + co;
      *reinterpret_cast<uint4*>(d) = hi;
      *reinterpret_cast<uint4*>(d + y_plane) = lo;
    }
  } else if (out_dtype == PPY_BF16) {
    const uint4 v = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
#pragma unroll
    for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(y) + rows[q] * y_ld + co) = v;
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + rows[q] * y_ld + co);
      dst[0] = a; dst[1] = b;
    }
  }
}

// Generic (unaligned / partial-vector / upsampling / fp32-residual) epilogue for 8 channels of one output row.
// Kept out of line so the hot path of the kernel stays small enough for the instruction cache.
__device__ __noinline__ void epilogue_slow(const ppy_conv_params& p, float4 va, float4 vb, int m, int co, int ho, int wo, float slope) {
  // (p: the CTA's shared-memory copy of the parameter block; the eight values by value -- registers, not a local array)
  const int ncol = (p.cout - co) < 8 ? (p.cout - co) : 8;
  const int hw_out = ho * wo;
  const int pix = m % hw_out, img = m / hw_out;
  const int oy = pix / wo, ox = pix % wo;
  for (int e = 0; e < ncol; ++e) {
    float f = e == 0 ? va.x : e == 1 ? va.y : e == 2 ? va.z : e == 3 ? va.w : e == 4 ? vb.x : e == 5 ? vb.y : e == 6 ? vb.z : vb.w;
    if (p.bias_map) f += __ldg(p.bias_map + (size_t)pix * p.cout + co + e);
    if (p.coord_w) f += __ldg(p.coord_w + co + e) * (__fdiv_rn((float)ox, (float)(wo - 1)) * 2.f - 1.f) +
                        __ldg(p.coord_w + p.cout + co + e) * (__fdiv_rn((float)oy, (float)(ho - 1)) * 2.f - 1.f);
    f = f * __ldg(p.scale + co + e) + __ldg(p.shift + co + e);
    if (p.residual) {
      if (p.out_dtype == PPY_F16X2) {
        const __half* r = reinterpret_cast<const __half*>(p.residual) + (size_t)m * p.res_ld + co + e;
        f += __half2float(r[0]) + __half2float(r[p.res_plane]);
      } else if (p.out_dtype == PPY_BF16) f += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.residual)[(size_t)m * p.res_ld + co + e]);
      else f += reinterpret_cast<const float*>(p.residual)[(size_t)m * p.res_ld + co + e];
    }
    f = f > 0.f ? f : f * slope;
    const int reps = p.upsample2x ? 4 : 1;
    for (int q = 0; q < reps; ++q) {
      size_t drow = (size_t)m;
      if (p.upsample2x) drow = ((size_t)img * 2 * ho + 2 * oy + (q >> 1)) * 2 * wo + 2 * ox + (q & 1);
      if (p.out_dtype == PPY_F16X2) {
        __half* d = reinterpret_cast<__half*>(p.y) + drow * p.y_ld + co + e;
        const __half hi = __float2half_rn(f);
        d[0] = hi;
        d[p.y_plane] = __float2half_rn(f - __half2float(hi));
        if (!(fabsf(f) <= 65504.f) && p.overflow) *p.overflow = 1;
      } else if (p.out_dtype == PPY_BF16) reinterpret_cast<__nv_bfloat16*>(p.y)[drow * p.y_ld + co + e] = __float2bfloat16_rn(f);
      else reinterpret_cast<float*>(p.y)[drow * p.y_ld + co + e] = f;
    }
  }
}

// Work unit of the persistent loop: (tap, M tile, N tile, K split); N tile fastest among the tiles that share an A tile,
// K splits of one tile adjacent.  Plain convs have one tap and one split, so unit == tile (n fastest) as before.
struct Unit { int mt, nt, tap, sp, kb0, kb1; };
template <bool ACC>
__device__ __forceinline__ Unit decode_unit(int u, int num_m_tiles, int num_n_tiles, int num_splits, int num_kb) {
  Unit r;
  if (!ACC) { r.sp = 0; r.tap = 0; r.nt = u % num_n_tiles; r.mt = u / num_n_tiles; r.kb0 = 0; r.kb1 = num_kb; return r; }
  r.sp = u % num_splits;
  int t = u / num_splits;
  r.nt = t % num_n_tiles; t /= num_n_tiles;
  r.mt = t % num_m_tiles;
  r.tap = t / num_m_tiles;
  const int per = (num_kb + num_splits - 1) / num_splits;
  r.kb0 = r.sp * per;
  r.kb1 = r.kb0 + per < num_kb ? r.kb0 + per : num_kb;
  return r;
}

// Partial-sum epilogue (K splits / weight-gradient taps): 8 channels of one output row are ADDED to the caller-zeroed fp32
// output with red.global; y_col = first output column of this tap, the shift is contributed by the first split only.
// (every argument by value: a `const ppy_conv_params&` forces a copy of the parameter block into LOCAL memory, and with 200+ KB of
// shared memory carved out of the L1 each of its loads -- and of the accumulator array passed by pointer -- was an L2 round trip:
// ncu showed ~2,400 cycles per call, ~30 us per launch whatever the GEMM's size)
__device__ __forceinline__ void epilogue_acc(float* __restrict__ y, int y_ld, int cout, const float* __restrict__ scale,
                                             const float* __restrict__ shift, float4 a, float4 b, int m, int co, int y_col, bool first_split) {
  const int ncol = (cout - co) < 8 ? (cout - co) : 8;
  float* dst = y + (size_t)m * y_ld + y_col + co;
  if (ncol == 8 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (reinterpret_cast<uintptr_t>(scale + co) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(shift + co) & 15) == 0) {
    // two 16-byte vector reductions instead of eight scalar ones: the L2 atomic units see a quarter of the requests
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + co)), s1 = __ldg(reinterpret_cast<const float4*>(scale + co) + 1);
    float4 h0 = make_float4(0.f, 0.f, 0.f, 0.f), h1 = h0;
    if (first_split) { h0 = __ldg(reinterpret_cast<const float4*>(shift + co)); h1 = __ldg(reinterpret_cast<const float4*>(shift + co) + 1); }
    atomicAdd(reinterpret_cast<float4*>(dst), make_float4(a.x * s0.x + h0.x, a.y * s0.y + h0.y, a.z * s0.z + h0.z, a.w * s0.w + h0.w));
    atomicAdd(reinterpret_cast<float4*>(dst) + 1, make_float4(b.x * s1.x + h1.x, b.y * s1.y + h1.y, b.z * s1.z + h1.z, b.w * s1.w + h1.w));
    return;
  }
  const float acc8[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  const unsigned long long g = (unsigned long long)__cvta_generic_to_global(dst);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (e < ncol) {
      const float f = acc8[e] * __ldg(scale + co + e) + (first_split ? __ldg(shift + co + e) : 0.f);
      asm volatile("red.global.add.f32 [%0], %1;" ::"l"(g + 4ull * e), "f"(f) : "memory");
    }
  }
}

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
template <int BN, int MODE, int EPI, bool ACC, bool CTA2, bool CHUNKED>
__global__ void __launch_bounds__(num_threads(MODE), 1)
conv_umma_kernel(const ppy_conv_params p, const int ho, const int wo, const int num_kb, const int num_m_tiles,
                 const int num_n_tiles, const int num_splits, const int num_taps, const int pw_tiles, const int ph_tiles,
                 const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_y,
                 const __grid_constant__ CUtensorMap tmap_r, const __grid_constant__ CUtensorMap tmap_a2 /* split: lo plane of A */) {
  // CTA2: the kernel runs as clusters of two CTAs (one TPC); the pair shares a 256 x BN tile -- CTA `cta_rank` stages and drains
  // M tile 2*unit + cta_rank and stages B rows [rank*BN/2, +BN/2); the leader (rank 0) issues tcgen05.mma.cta_group::2 over both
  // CTAs' shared memory, its commits multicast to both CTAs' barriers.  num_m_tiles then counts PAIRS of M tiles.
  static_assert(!CTA2 || (mode_is_tma(MODE) && !ACC), "the CTA-pair kernel is TMA-fed only");
  static_assert(!SPLIT || (EPI == EPI_SLAB && !ACC), "split mode: slab epilogue, no partial sums");
  static_assert(!CHUNKED || (SPLIT && BN <= 128), "K chunks: split mode, at most two sub-tiles per epilogue warp");
  using Cfg = TileCfg<BN, MODE, EPI, CTA2>;
  constexpr int S = Cfg::kStages;
  constexpr int A_STAGE = Cfg::kAStageBytes;
  constexpr int A_TILE = Cfg::kATileBytes, B_PLANE = Cfg::kBPlaneBytes;      // split mode: lo tiles follow the hi tiles of a stage
  constexpr int TMA_WARP = tma_warp(MODE), MMA_WARP = mma_warp(MODE);
  constexpr int PW = patch_w(MODE), PH = patch_h(MODE);    // pixel patch of a tile (MODE_TMA_PATCH / MODE_TMA_SLAB)
  constexpr int CP_LAG = Cfg::kCpLag;
  const int cta_rank = CTA2 ? (int)(blockIdx.x & 1) : 0;
  const int tile_first = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto tile_mt = [&](int tile) { return CTA2 ? 2 * (tile / num_n_tiles) + cta_rank : tile / num_n_tiles; };
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + S * A_STAGE;
  const uint32_t stg_off = S * (A_STAGE + Cfg::kBStageBytes);
  // barriers: full[S], empty[S], tmem_full[2], tmem_empty[2], res_full[8 warps][2], then the TMEM base slot
  const uint32_t bars = smem_base + stg_off + Cfg::kEpiBytes;
  // Accumulator buffers in TMEM: two (tile i drains while tile i+1 is multiplied); K-chunked tiles use four, so the MMA issuer can
  // run up to four chunks ahead while the epilogue warps are still busy with the previous tile's coalesced pass (with two, layers
  // with few chunks per tile -- K = 1152 -- stalled the tensor pipe for ~25 % of the time)
  constexpr int NBUF = CHUNKED ? 4 : 2;
  constexpr int TMEM_COLS = NBUF * BN < 32 ? 32 : NBUF * BN;
  static_assert(TMEM_COLS <= 512, "TMEM budget");
  volatile uint32_t* tmem_ptr_slot = reinterpret_cast<volatile uint32_t*>(gen_base + stg_off + Cfg::kEpiBytes + (2 * S + 8 + 2 * EPI_WARPS) * 8);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
  auto tmem_full_bar = [&](int a) { return bars + 8u * (2 * S + a); };
  auto tmem_empty_bar = [&](int a) { return bars + 8u * (2 * S + 4 + a); };
  auto res_full_bar = [&](int w, int b) { return bars + 8u * (2 * S + 8 + 2 * w + b); };

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform for the compiler: role branches are not divergent
  // The out-of-line generic epilogue takes the parameter block by reference.  Handing it the kernel parameter itself makes the
  // compiler keep a per-thread copy in LOCAL memory -- and next to 200+ KB of shared memory the L1 holds nothing, so every field
  // access there is an L2 round trip (~2,400 cycles per call: the 2-channel tail convs of the head outputs and the 27-channel DCN
  // offset convs take that path for every row).  One copy per CTA in shared memory instead.
  static_assert(sizeof(ppy_conv_params) <= Cfg::kParamBytes, "parameter block larger than its shared-memory slot");
  ppy_conv_params* const p_smem = reinterpret_cast<ppy_conv_params*>(gen_base + stg_off + Cfg::kEpiBytes + 512);
  if (tid == 32) *p_smem = p;                                      // (published by the barrier below)
  const long long M = (long long)p.n * ho * wo;
  const int num_tiles = ACC ? num_m_tiles * num_n_tiles * num_splits * num_taps : num_m_tiles * num_n_tiles;

  if (warp == TMA_WARP && lane == 0) {
    // the copy engine fetches a tensor map (128 B in the kernel parameter space) on first use: start those fetches now, under
    // the prologue and the tail of the previous kernel, instead of in front of the first operand load
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
    if (mode_is_tma(MODE)) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
    if (mode_is_tma(MODE) && SPLIT) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a2)) : "memory");
    if (EPI == EPI_TMA) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_y)) : "memory");
      if (p.residual) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_r)) : "memory");
    }
  }
  if (tid == 0) {
    const uint32_t full_count = mode_is_tma(MODE) ? 1u : (uint32_t)(BLOCK_M + 1);
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), full_count); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < NBUF; ++a) {
      // TMA epilogue at BLOCK_N 64 (one column group): the two warps of a TMEM lane quarter take alternate tiles, so each
      // accumulator buffer is drained by four warps
      // (the slab epilogue at BLOCK_N 32 -- one sub-tile per tile -- does the same)
      constexpr uint32_t drainers = ((EPI == EPI_TMA && BN == GROUP_COLS) || (EPI == EPI_SLAB && BN == SUB && !CHUNKED && !ACC)) ? EPI_WARPS / 2 : EPI_WARPS;
      mbar_init(tmem_full_bar(a), 1); mbar_init(tmem_empty_bar(a), CTA2 ? 2 * drainers : drainers);
    }
    for (int w = 0; w < EPI_WARPS; ++w) { mbar_init(res_full_bar(w, 0), 1); mbar_init(res_full_bar(w, 1), 1); }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) {
    if (CTA2) tmem_alloc2(smem_u32(const_cast<uint32_t*>(tmem_ptr_slot)), TMEM_COLS);
    else tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_slot)), TMEM_COLS);
  }
  tc_fence_before();
  if (CTA2) cluster_sync_all(); else __syncthreads();      // pair: the peer's barriers are initialised before anything targets them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation) may overlap the tail of the previous kernel
  // of the stream; nothing below touches global memory before that kernel has completed and its writes are visible.  The
  // dependents of THIS grid may be scheduled as soon as its CTAs free their SMs.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (!mode_is_tma(MODE) && warp >= PRODUCER_WARP0 && warp < PRODUCER_WARP0 + 4) {
    // =====================================================================================
    // A producers
    // =====================================================================================
    const int ptid = tid - PRODUCER_WARP0 * 32;
    const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(p.x);
    const int taps = p.kh * p.kw;
    if (MODE == MODE_GATHER) {
      const int j = ptid & 7, rg = ptid >> 3;
      const uint32_t dst0 = (uint32_t)rg * 128u + (((uint32_t)j ^ (uint32_t)(rg & 7)) << 4);
      int g = 0;                                   // global K-block counter (ring position)
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const Unit u = decode_unit<ACC>(tile, num_m_tiles, num_n_tiles, num_splits, num_kb);
        const long long m0 = (long long)u.mt * BLOCK_M;
        int iy0[8], ix0[8];
        long long pbase[8];                        // element offset of pixel (img, 0, 0); < 0 = row beyond M
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const long long m = m0 + rg + 16 * i;
          if (m < M) {
            const unsigned mu = (unsigned)m, pix = mu % (unsigned)(wo * ho), img = mu / (unsigned)(wo * ho);
            const int oy = (int)(pix / (unsigned)wo), ox = (int)(pix % (unsigned)wo);
            iy0[i] = oy * p.stride - p.pad; ix0[i] = ox * p.stride - p.pad;
            pbase[i] = (long long)img * p.h * p.w;
          } else { iy0[i] = 0; ix0[i] = 0; pbase[i] = -1; }
        }
        int tap = 0, c = j * 8 + u.kb0 * BLOCK_K, ky = 0, kx = 0;
        while (c >= p.cin) { c -= p.cin; ++tap; if (++kx == p.kw) { kx = 0; ++ky; } }
        for (int kb = u.kb0; kb < u.kb1; ++kb, ++g) {
          const int s = g % S;
          mbar_wait(empty_bar(s), ((g / S) & 1) ^ 1);
          const uint32_t dst = smem_a + s * A_STAGE + dst0;
          const bool tap_ok = tap < taps;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int iy = iy0[i] + ky, ix = ix0[i] + kx;
            const bool ok = tap_ok && pbase[i] >= 0 && iy >= 0 && iy < p.h && ix >= 0 && ix < p.w;
            const __nv_bfloat16* src = ok ? x + (pbase[i] + (long long)iy * p.w + ix) * p.x_ld + c : x;
            cp_async16(dst + (uint32_t)i * (16u * 128u), src, ok ? 16u : 0u);
            if (SPLIT) cp_async16(dst + (uint32_t)A_TILE + (uint32_t)i * (16u * 128u), ok ? src + p.x_plane : x, ok ? 16u : 0u);
          }
          cp_async_commit();
          c += BLOCK_K;
          while (c >= p.cin) { c -= p.cin; ++tap; if (++kx == p.kw) { kx = 0; ++ky; } }
          if (g >= CP_LAG) {
            cp_async_wait<CP_LAG>();
            fence_proxy_async();
            mbar_arrive(full_bar((g - CP_LAG) % S));
          }
        }
      }
      cp_async_wait<0>();
      fence_proxy_async();
      for (int q = (g > CP_LAG ? g - CP_LAG : 0); q < g; ++q) mbar_arrive(full_bar(q % S));
    } else if (MODE == MODE_DCN) {
      // thread = tile row; one tap per K block (cin % 64 == 0)
      const int r = ptid;
      const uint32_t row_off = (uint32_t)r * 128u;
      const uint32_t sw = (uint32_t)(r & 7);
      const int kb_per_tap = p.cin / BLOCK_K;
      int g = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const long long m = (long long)(tile / num_n_tiles) * BLOCK_M + r;
        const bool valid = m < M;
        int img = 0, oy = 0, ox = 0;
        if (valid) { ox = (int)(m % wo); oy = (int)((m / wo) % ho); img = (int)(m / ((long long)wo * ho)); }
        const float* om = valid ? p.offset_mask + m * p.om_ld : nullptr;
        float w4[4] = {0.f, 0.f, 0.f, 0.f};
        const __nv_bfloat16* src4[4] = {nullptr, nullptr, nullptr, nullptr};
        for (int kb = 0; kb < num_kb; ++kb, ++g) {
          const int s = g % S;
          const int tap = kb / kb_per_tap, c0 = (kb % kb_per_tap) * BLOCK_K;
          if (c0 == 0) {                          // new tap: sampling position, corner pointers and weights
#pragma unroll
            for (int q = 0; q < 4; ++q) { src4[q] = nullptr; w4[q] = 0.f; }
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
                  src4[q] = x + (((long long)img * p.h + yy) * p.w + xx) * p.x_ld;
                  w4[q] = wq[q] * mask;
                }
              }
            }
          }
          mbar_wait(empty_bar(s), ((g / S) & 1) ^ 1);
          const uint32_t dst_row = smem_a + s * A_STAGE + row_off;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (src4[q]) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(src4[q] + c0 + 8 * j));
                const float wq = w4[q];
                if (SPLIT) {                       // corner value = hi + lo (exact in fp32)
                  const uint4 v2 = __ldg(reinterpret_cast<const uint4*>(src4[q] + p.x_plane + c0 + 8 * j));
                  float2 t, u;
                  t = h2f(v.x); u = h2f(v2.x); acc[0] += wq * (t.x + u.x); acc[1] += wq * (t.y + u.y);
                  t = h2f(v.y); u = h2f(v2.y); acc[2] += wq * (t.x + u.x); acc[3] += wq * (t.y + u.y);
                  t = h2f(v.z); u = h2f(v2.z); acc[4] += wq * (t.x + u.x); acc[5] += wq * (t.y + u.y);
                  t = h2f(v.w); u = h2f(v2.w); acc[6] += wq * (t.x + u.x); acc[7] += wq * (t.y + u.y);
                } else {
                acc[0] += wq * bf_lo(v.x); acc[1] += wq * bf_hi(v.x); acc[2] += wq * bf_lo(v.y); acc[3] += wq * bf_hi(v.y);
                acc[4] += wq * bf_lo(v.z); acc[5] += wq * bf_hi(v.z); acc[6] += wq * bf_lo(v.w); acc[7] += wq * bf_hi(v.w);
                }
              }
            }
            const uint32_t d = dst_row + (((uint32_t)j ^ sw) << 4);
            if (SPLIT) {
              uint4 hi, lo;
              split8(make_float4(acc[0], acc[1], acc[2], acc[3]), make_float4(acc[4], acc[5], acc[6], acc[7]), hi, lo, nullptr);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d + (uint32_t)A_TILE), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w) : "memory");
            } else
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d), "r"(pack_bf16(acc[0], acc[1])),
                         "r"(pack_bf16(acc[2], acc[3])), "r"(pack_bf16(acc[4], acc[5])), "r"(pack_bf16(acc[6], acc[7])) : "memory");
          }
          fence_proxy_async();
          mbar_arrive(full_bar(s));
        }
      }
    }
    // MODE_TMA_A: producers have nothing to do
  } else if (warp == TMA_WARP) {
    // =====================================================================================
    // TMA producer (weights; + activations in tma_a mode)
    // =====================================================================================
    {
      constexpr uint32_t tx_bytes = (Cfg::kBStageBytes + (mode_is_tma(MODE) ? A_STAGE : 0)) * (CTA2 ? 2 : 1);
      const int kb_per_tap = p.cin / BLOCK_K;
      // second K source (split mode, 1x1 convs; see ppy_conv_params.x2): K blocks >= kb_x2 read x2 through the maps tmap_y (hi
      // plane) / tmap_r (lo plane), which the slab epilogue leaves unused
      constexpr bool X2_MODE = SPLIT && (MODE == MODE_TMA_A || MODE == MODE_TMA_IM2COL || MODE == MODE_TMA_PATCH);
      const int kb_x2 = X2_MODE ? num_kb - p.x2_kb : num_kb;
      const bool x2_ident = SPLIT && MODE == MODE_TMA_A && p.x2_tiled != 0;      // identity blocks: their lo weight plane is zero
      int g = 0;
      for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
        const Unit u = decode_unit<ACC>(tile, num_m_tiles, num_n_tiles, num_splits, num_kb);
        const int n0 = u.nt * BN + (CTA2 ? cta_rank * (BN / 2) : 0);
        const int mt = CTA2 ? 2 * u.mt + cta_rank : u.mt;
        const int m0 = mt * BLOCK_M;
        // weight-gradient GEMM (p.wgrad_pitch > 0): the B operand is the transposed, zero-bordered activation read at the
        // flat pixel offset of this tap -- (ky-1)*pitch + (kx-1)
        const int b_shift = (ACC && num_taps == 9) ? (u.tap / 3 - 1) * p.wgrad_pitch + (u.tap % 3 - 1) : 0;
        int px0 = 0, py0 = 0, img = 0;
        if (mode_is_patchy(MODE)) {
          px0 = (mt % pw_tiles) * PW; py0 = ((mt / pw_tiles) % ph_tiles) * PH; img = mt / (pw_tiles * ph_tiles);
        }
        if (MODE == MODE_TMA_IM2COL) {           // base pixel of the tile's first output pixel
          const unsigned hw_out = (unsigned)(ho * wo), pix = (unsigned)m0 % hw_out;
          img = (int)((unsigned)m0 / hw_out);
          px0 = (int)(pix % (unsigned)wo) * p.stride - p.pad; py0 = (int)(pix / (unsigned)wo) * p.stride - p.pad;
        }
        for (int kb = u.kb0; kb < u.kb1; ++kb, ++g) {
          const int s = g % S;
          mbar_wait(empty_bar(s), ((g / S) & 1) ^ 1);
          if (elect_one()) {                     // the whole warp walks the ring (uniform control flow), one lane issues
            // CTA pairs: both CTAs' bytes are counted on the leader's barrier, which the leader arms for the whole pair
            const uint32_t bar = CTA2 ? map_to_cta(full_bar(s), 0) : full_bar(s);
            const bool second = kb >= kb_x2, skip_blo = second && x2_ident;
            if (!CTA2 || cta_rank == 0) mbar_arrive_expect_tx(full_bar(s), tx_bytes - (skip_blo ? (uint32_t)B_PLANE * (CTA2 ? 2u : 1u) : 0u));
            constexpr bool by_tap = MODE == MODE_TMA_PATCH || MODE == MODE_TMA_IM2COL;
            const int tap = by_tap ? kb / kb_per_tap : 0, c0 = by_tap ? (kb % kb_per_tap) * BLOCK_K : 0;
#pragma unroll
            for (int pl = 0; pl < PLANES; ++pl) {    // split mode: the hi tiles, then the lo tiles (second A map; weight rows + cout_pad)
              const CUtensorMap* ma = pl ? &tmap_a2 : &tmap_a;
              const uint32_t a_dst = smem_a + s * A_STAGE + pl * A_TILE, b_dst = smem_b + s * Cfg::kBStageBytes + pl * B_PLANE;
              const int nrow = n0 + pl * p.cout_pad;
              if (X2_MODE && second) {                  // block of the second K source: maps tmap_y (hi plane) / tmap_r (lo plane)
                const CUtensorMap* m2 = pl ? &tmap_r : &tmap_y;
                const int col = (kb - kb_x2) * BLOCK_K + (x2_ident ? u.nt * BN : 0);
                if (MODE == MODE_TMA_PATCH) {           // (batch-invariant sources have one image)
                  const int img2 = p.x2_row_mod ? 0 : img;
                  if (CTA2) tma2_load_4d(a_dst, m2, bar, col, px0, py0, img2); else tma_load_4d(a_dst, m2, bar, col, px0, py0, img2);
                } else {
                  const int row = p.x2_row_mod ? m0 % p.x2_row_mod : m0;
                  if (CTA2) tma2_load_2d(a_dst, m2, bar, col, row); else tma_load_2d(a_dst, m2, bar, col, row);
                }
                if (pl == 1 && skip_blo) continue;       // identity block: no lo weight tile
              } else {
              if (MODE == MODE_TMA_A) { if (CTA2) tma2_load_2d(a_dst, ma, bar, kb * BLOCK_K, m0); else tma_load_2d(a_dst, ma, bar, kb * BLOCK_K, m0); }
              if (MODE == MODE_TMA_PATCH) {
                // one tap x 64 channels of the 16x8 patch; the halo (negative / beyond-edge coordinates) is zero-filled by TMA
                if (CTA2) tma2_load_4d(a_dst, ma, bar, c0, px0 + tap % 3 - 1, py0 + tap / 3 - 1, img);
                else tma_load_4d(a_dst, ma, bar, c0, px0 + tap % 3 - 1, py0 + tap / 3 - 1, img);
              }
              if (MODE == MODE_TMA_IM2COL) {
                if (CTA2) tma2_load_im2col(a_dst, ma, bar, c0, px0, py0, img, (uint16_t)(tap % p.kw), (uint16_t)(tap / p.kw));
                else tma_load_im2col(a_dst, ma, bar, c0, px0, py0, img, (uint16_t)(tap % p.kw), (uint16_t)(tap / p.kw));
              }
              }
              if (MODE == MODE_TMA_SLAB) {           // iteration kb = (channel block, kx): the slab and the weight tiles of its three taps
                const int cb = kb / 3, kx = kb % 3;
                if (CTA2) tma2_load_4d(a_dst, ma, bar, cb * BLOCK_K, px0 + kx - 1, py0 - 1, img);
                else tma_load_4d(a_dst, ma, bar, cb * BLOCK_K, px0 + kx - 1, py0 - 1, img);
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                  if (CTA2) tma2_load_2d(b_dst + ky * Cfg::kBTileBytes, &tmap_b, bar, (ky * 3 + kx) * p.cin + cb * BLOCK_K, nrow);
                  else tma_load_2d(b_dst + ky * Cfg::kBTileBytes, &tmap_b, bar, (ky * 3 + kx) * p.cin + cb * BLOCK_K, nrow);
                }
              } else {
                if (CTA2) tma2_load_2d(b_dst, &tmap_b, bar, kb * BLOCK_K, nrow);
                else tma_load_2d(b_dst, &tmap_b, bar, kb * BLOCK_K + b_shift, nrow);
              }
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // =====================================================================================
    // MMA issuer
    // =====================================================================================
    if (cta_rank == 0) {
      constexpr uint32_t idesc = make_idesc(CTA2 ? 2 * BLOCK_M : BLOCK_M, BN);
      constexpr int CH = CHUNKED ? chunk_kb(MODE) : (1 << 28);
      const int kb_ident = (SPLIT && MODE == MODE_TMA_A && p.x2_tiled) ? num_kb - p.x2_kb : num_kb;   // identity blocks: b_lo == 0
      int g = 0, it = 0;                         // it: uses of the accumulator buffers = tiles (CHUNKED: K chunks)
      for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
        const Unit u = decode_unit<ACC>(tile, num_m_tiles, num_n_tiles, num_splits, num_kb);
        for (int kc0 = u.kb0; kc0 < u.kb1; kc0 += CH, ++it) {
        const int acc = it % NBUF;
        mbar_wait(tmem_empty_bar(acc), ((it / NBUF) & 1) ^ 1);      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        const int kc1 = (u.kb1 - kc0 > CH) ? kc0 + CH : u.kb1;
        for (int kb = kc0; kb < kc1; ++kb, ++g) {
          const int s = g % S;
          mbar_wait(full_bar(s), (g / S) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_a + s * A_STAGE, b_addr = smem_b + s * Cfg::kBStageBytes;
          if (elect_one()) {                     // warp-uniform loop, one lane issues the MMAs and their commit
          // split mode: three operand combinations per K block -- hi*hi, hi*lo, lo*hi -- into the same accumulator
          constexpr int COMBOS = SPLIT ? 3 : 1;
#pragma unroll
          for (int cb = 0; cb < COMBOS; ++cb) {
          if (SPLIT && cb == 1 && kb >= kb_ident) continue;
          const uint32_t a_pl = a_addr + (cb == 2 ? A_TILE : 0), b_pl = b_addr + (cb == 1 ? B_PLANE : 0);
          if (MODE == MODE_TMA_SLAB) {           // three taps per stage: tap ky reads the slab one pixel row (1024 bytes) further down
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
              for (int k = 0; k < BLOCK_K / 16; ++k) {
                const uint64_t da = mk_desc(a_pl + ky * ATOM_BYTES + k * 32), db = mk_desc(b_pl + ky * Cfg::kBTileBytes + k * 32);
                if (CTA2) umma2_bf16(d_tmem, da, db, idesc, ((kb - kc0) | cb | ky | k) ? 1u : 0u);
                else umma_bf16(d_tmem, da, db, idesc, ((kb - kc0) | cb | ky | k) ? 1u : 0u);
              }
            }
          } else {
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            if (CTA2) umma2_bf16(d_tmem, mk_desc(a_pl + k * 32), mk_desc(b_pl + k * 32), idesc, ((kb - kc0) | cb | k) ? 1u : 0u);
            else umma_bf16(d_tmem, mk_desc(a_pl + k * 32), mk_desc(b_pl + k * 32), idesc, ((kb - kc0) | cb | k) ? 1u : 0u);
          }
          }
          }
          if (CTA2) umma_commit2(empty_bar(s)); else umma_commit(empty_bar(s));   // frees the stage (in both CTAs) once read
          }
          __syncwarp();
        }
        if (elect_one()) {
          if (CTA2) umma_commit2(tmem_full_bar(acc)); else umma_commit(tmem_full_bar(acc));   // accumulator (chunk) complete -> epilogue(s)
        }
        __syncwarp();
        }
      }
    }
  } else if (EPI == EPI_TMA) {
    // =====================================================================================
    // EPI_TMA epilogue, warps 0-7: no global-memory instruction and no CTA-wide barrier.  Every warp works alone on
    // [32 rows x 64 columns] boxes of the tile -- warp w owns TMEM lanes / tile rows 32*(w&3).. and the 64-column groups
    // g = (w>>2), (w>>2)+2, ..: its lane 0 fetches the residual box by TMA two boxes ahead (double buffered), every lane (= one
    // tile row) turns 64 accumulator columns into bf16 with the CoordConv term, scale/shift, residual and activation, writes
    // them into the swizzled output box, and lane 0 hands the box to a TMA store (rows beyond M / columns beyond cout are
    // clipped by the copy engine, so there are no masks here).  Warps drift apart freely, overlapping each other's latencies.
    // =====================================================================================
    constexpr int G = BN / GROUP_COLS;
    constexpr int GPW = G >= 2 ? G / 2 : 1;                 // groups per warp and tile
    constexpr int RB = Cfg::kResBufs;
    // G >= 2: the two warps of a TMEM lane quarter split the tile's column groups.  G == 1: they take alternate TILES (warp
    // half h drains accumulator buffer h), so all eight warps stay busy without sharing a box.
    constexpr int TSTRIDE = G >= 2 ? 1 : 2;
    const int quarter = warp & 3, half = warp >> 2;
    const int my_first = G >= 2 ? tile_first : tile_first + half * tile_step, my_step = TSTRIDE * tile_step;
    const uint32_t res_s = smem_base + stg_off + (uint32_t)warp * Cfg::kEpiWarpBytes, out_s = res_s + RB * BOX_BYTES;
    float* tab = reinterpret_cast<float*>(gen_base + stg_off + EPI_WARPS * Cfg::kEpiWarpBytes) + warp * EPI_TABLE_FLOATS;
    const float slope = p.act == PPY_ACT_RELU ? 0.f : (p.act == PPY_ACT_LEAKY ? 0.1f : 1.f);
    const bool has_res = p.residual != nullptr, has_coord = p.coord_w != nullptr;
    const bool out_f32 = p.out_dtype == PPY_F32;           // fp32 output: the group leaves as two [32 x 32 fp32] boxes (no residual)
    const unsigned hw_out = (unsigned)(ho * wo);
    const int row = quarter * 32 + lane;                    // tile row of this lane
    const uint32_t row_off = (uint32_t)lane * 128u;
    const uint32_t sw = (uint32_t)(lane & 7);
    const uint32_t empty_rank0 = CTA2 ? map_to_cta(tmem_empty_bar(0), 0) : 0u;      // the leader's MMA thread waits for both epilogues
    // box j of this warp: tile = my_first + (j / GPW) * my_step, group = half + 2 * (j % GPW) (G >= 2) or 0
    auto box_col0 = [&](int tile_, int gi) { return (tile_ % num_n_tiles) * BN + (G >= 2 ? half + 2 * gi : 0) * GROUP_COLS; };
    auto box_move = [&](bool load, uint32_t smem, uint32_t bar, int tile_, int col0) {      // one lane: residual load / output store
      const int mt_ = tile_mt(tile_);
      if (mode_is_patchy(MODE)) {
        const int x0 = (mt_ % pw_tiles) * PW, y0 = ((mt_ / pw_tiles) % ph_tiles) * PH + quarter * (BOX_ROWS / PW), img = mt_ / (pw_tiles * ph_tiles);
        if (load) tma_load_4d(smem, &tmap_r, bar, col0, x0, y0, img); else tma_store_4d(&tmap_y, smem, col0, x0, y0, img);
      } else {
        if (load) tma_load_2d(smem, &tmap_r, bar, col0, mt_ * BLOCK_M + quarter * BOX_ROWS); else tma_store_2d(&tmap_y, smem, col0, mt_ * BLOCK_M + quarter * BOX_ROWS);
      }
    };
    int ji = 0, li = 0, lc = 0;                             // next box to request; live boxes requested / consumed
    auto request_next = [&]() {                             // residual prefetch: skip boxes beyond cout, stop at the end of the walk
      for (;;) {
        const int tile_ = my_first + (ji / GPW) * my_step;
        if (tile_ >= num_tiles) return;
        const int col0 = box_col0(tile_, ji % GPW);
        ++ji;
        if (col0 < p.cout) {
          if (lane == 0) {
            mbar_arrive_expect_tx(res_full_bar(warp, li % RB), BOX_BYTES);
            box_move(true, res_s + (li % RB) * BOX_BYTES, res_full_bar(warp, li % RB), tile_, col0);
          }
          ++li;
          return;
        }
      }
    };
    if (has_res) { request_next(); if (RB == 2) request_next(); }
    int it = G >= 2 ? 0 : half, tab_col0 = -1;             // `it` counts the CTA's tiles (accumulator buffer / phase bookkeeping)
    int nbox = 0;                                           // boxes stored by this warp
    for (int tile = my_first; tile < num_tiles; tile += my_step, it += TSTRIDE) {
      const int acc = it & 1;
      const int mt = tile_mt(tile);
      float xc = 0.f, yc = 0.f;                             // CoordConv coordinates of this lane's output pixel
      if (has_coord) {
        const unsigned pix = (unsigned)(mt * BLOCK_M + row) % hw_out;      // coord_w convs are 1x1: linear M tiles
        xc = __fdiv_rn((float)(pix % (unsigned)wo), (float)(wo - 1)) * 2.f - 1.f;
        yc = __fdiv_rn((float)(pix / (unsigned)wo), (float)(ho - 1)) * 2.f - 1.f;
      }
      mbar_wait(tmem_full_bar(acc), (it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
      bool released = false;
      auto release_acc = [&]() {                            // this warp's TMEM reads of the tile are done
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (CTA2) mbar_arrive_cluster(empty_rank0 + 8u * acc); else mbar_arrive(tmem_empty_bar(acc)); }
        released = true;
      };
      {
#pragma unroll 1
        for (int gi = 0; gi < GPW; ++gi) {
          const int col0 = box_col0(tile, gi);
          if (col0 >= p.cout) continue;                     // whole group beyond cout
          const int gcol = col0 - (tile % num_n_tiles) * BN;     // column of the group inside the accumulator
          uint32_t v[64];
          tmem_ld32_nowait(t_row + (uint32_t)gcol, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
          tmem_ld32_nowait(t_row + (uint32_t)(gcol + 32), *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
          if (col0 != tab_col0) {                           // (re)load this warp's per-channel table: lane -> two columns
            __syncwarp();
            const int c = col0 + 2 * lane;
            const bool ok0 = c < p.cout, ok1 = c + 1 < p.cout;
            tab[2 * lane] = ok0 ? __ldg(p.scale + c) : 0.f;                 tab[2 * lane + 1] = ok1 ? __ldg(p.scale + c + 1) : 0.f;
            tab[64 + 2 * lane] = ok0 ? __ldg(p.shift + c) : 0.f;            tab[64 + 2 * lane + 1] = ok1 ? __ldg(p.shift + c + 1) : 0.f;
            if (has_coord) {
              tab[128 + 2 * lane] = ok0 ? __ldg(p.coord_w + c) : 0.f;          tab[128 + 2 * lane + 1] = ok1 ? __ldg(p.coord_w + c + 1) : 0.f;
              tab[192 + 2 * lane] = ok0 ? __ldg(p.coord_w + p.cout + c) : 0.f; tab[192 + 2 * lane + 1] = ok1 ? __ldg(p.coord_w + p.cout + c + 1) : 0.f;
            }
            tab_col0 = col0;
          }
          // output box rotation: with a residual the warp has ONE output box (wait for its previous store to drain); without,
          // the residual boxes serve as output boxes too (RB + 1 in rotation: only the store RB boxes back must have drained)
          const uint32_t out_box = (has_res || out_f32) ? (out_f32 ? res_s : out_s) : res_s + (uint32_t)(nbox % (RB + 1)) * BOX_BYTES;
          if (lane == 0) { if (has_res || out_f32) bulk_wait_read<0>(); else bulk_wait_read<RB>(); }
          if (has_res) mbar_wait(res_full_bar(warp, lc % RB), (lc / RB) & 1);
          tmem_wait_ld();
          __syncwarp();
          if (gi == GPW - 1) release_acc();
          const uint32_t rbase = res_s + (lc % RB) * BOX_BYTES + row_off, obase = out_box + row_off;
#pragma unroll
          for (int q = 0; q < 8; ++q) {                      // 8 channels (one 16-byte chunk) at a time
            const uint32_t chunk = ((uint32_t)q ^ sw) << 4;
            uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0;
            if (has_res) asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(rbase + chunk));
            const float4 s0 = *reinterpret_cast<const float4*>(tab + q * 8), s1 = *reinterpret_cast<const float4*>(tab + q * 8 + 4);
            const float4 h0 = *reinterpret_cast<const float4*>(tab + 64 + q * 8), h1 = *reinterpret_cast<const float4*>(tab + 64 + q * 8 + 4);
            if (has_coord) {                                 // rank-2 CoordConv term, added to the accumulator before scale/shift
              const float* cx = tab + 128 + q * 8;
              const float* cy = tab + 192 + q * 8;
#pragma unroll
              for (int e = 0; e < 8; ++e) v[8 * q + e] = __float_as_uint(__uint_as_float(v[8 * q + e]) + cx[e] * xc + cy[e] * yc);
            }
            float f[8];
            f[0] = __uint_as_float(v[8 * q + 0]) * s0.x + h0.x + bf_lo(r0); f[1] = __uint_as_float(v[8 * q + 1]) * s0.y + h0.y + bf_hi(r0);
            f[2] = __uint_as_float(v[8 * q + 2]) * s0.z + h0.z + bf_lo(r1); f[3] = __uint_as_float(v[8 * q + 3]) * s0.w + h0.w + bf_hi(r1);
            f[4] = __uint_as_float(v[8 * q + 4]) * s1.x + h1.x + bf_lo(r2); f[5] = __uint_as_float(v[8 * q + 5]) * s1.y + h1.y + bf_hi(r2);
            f[6] = __uint_as_float(v[8 * q + 6]) * s1.z + h1.z + bf_lo(r3); f[7] = __uint_as_float(v[8 * q + 7]) * s1.w + h1.w + bf_hi(r3);
            if (slope != 1.f) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], f[e] * slope);
            }
            if (out_f32) {                                   // channels 8q..8q+7 = 16-byte chunks 2(q&3), 2(q&3)+1 of box q>>2
              const uint32_t fb = obase + (uint32_t)(q >> 2) * BOX_BYTES;
              const uint32_t c0 = (((uint32_t)(q & 3) * 2u) ^ sw) << 4, c1 = (((uint32_t)(q & 3) * 2u + 1u) ^ sw) << 4;
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(fb + c0), "r"(__float_as_uint(f[0])), "r"(__float_as_uint(f[1])),
                           "r"(__float_as_uint(f[2])), "r"(__float_as_uint(f[3])) : "memory");
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(fb + c1), "r"(__float_as_uint(f[4])), "r"(__float_as_uint(f[5])),
                           "r"(__float_as_uint(f[6])), "r"(__float_as_uint(f[7])) : "memory");
            } else
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(obase + chunk), "r"(pack_bf16(f[0], f[1])),
                         "r"(pack_bf16(f[2], f[3])), "r"(pack_bf16(f[4], f[5])), "r"(pack_bf16(f[6], f[7])) : "memory");
          }
          fence_proxy_async();                              // generic-proxy accesses of both boxes -> ordered before the TMA store / reload
          __syncwarp();
          if (lane == 0) {
            box_move(false, out_box, 0u, tile, col0);
            if (out_f32 && col0 + GROUP_COLS / 2 < p.cout) box_move(false, out_box + BOX_BYTES, 0u, tile, col0 + GROUP_COLS / 2);
            bulk_commit();
          }
          ++nbox;
          if (has_res) { ++lc; request_next(); }            // the residual buffer just read is free: fetch the box two ahead
        }
      }
      if (!released) release_acc();
    }
    if (lane == 0) bulk_wait<0>();                          // this warp's output boxes are in global memory before the CTA exits
  } else {
    // =====================================================================================
    // epilogue warps 0-7: warp w reads TMEM lanes 32*(w&3).. and handles sub-tiles cc = (w>>2), (w>>2)+2, ...
    // =====================================================================================
    const int quarter = warp & 3, half = warp >> 2;
    float* slab = reinterpret_cast<float*>(gen_base + stg_off) + (size_t)warp * 32 * ST_LD;   // warp-private 32 x 32 fp32
    const uint32_t slab_u32 = smem_base + stg_off + (uint32_t)(warp * 32 * ST_LD * 4);
    const bool out_bf16 = p.out_dtype != PPY_F32;      // 16-bit output elements: bf16, or (split mode) the two fp16 planes of a pair
    const int esz = out_bf16 ? 2 : 4;
    const unsigned hw_out = (unsigned)(ho * wo);
    const int cpair = (lane & 3) * 2;            // this lane's two 16-byte chunks (8 channels) of a 32-column row
    const int colv = cpair * 4;
    const int rsub = lane >> 2;                  // 8 rows per pass, 4 passes
    const float slope = p.act == PPY_ACT_RELU ? 0.f : (p.act == PPY_ACT_LEAKY ? 0.1f : 1.f);
    // the fast path needs 16-byte aligned full vectors everywhere; anything else goes through epilogue_slow
    const bool aligned = ((p.y_ld * esz) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 15) == 0 &&
                         (!p.residual || (out_bf16 && ((p.res_ld * 2) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0)) &&
                         (!p.bias_map || (p.cout & 7) == 0) && !ACC && !p.coord_w &&   // partial sums (atomics), coord fold: out of line
                         (!SPLIT || (((p.y_plane * 2) & 15) == 0 && ((p.res_plane * 2) & 15) == 0));
    const bool has_res = p.residual != nullptr;
    constexpr int NSUB = BN / SUB;               // sub-tiles per tile
    constexpr int MY_SUBS = (NSUB + 1) / 2;      // upper bound of sub-tiles per warp
    // BLOCK_N 32 (one sub-tile per tile): the two warps of a TMEM lane quarter take alternate TILES -- warp half h drains
    // accumulator buffer h -- so all eight warps work (these layers are epilogue-bound: tiny N, short K)
    constexpr bool ALT = NSUB == 1 && !CHUNKED && !ACC;
    const int sub0 = ALT ? 0 : half;
    const int my_first = ALT ? tile_first + half * tile_step : tile_first, my_step = ALT ? 2 * tile_step : tile_step;
    auto row_to_m = [&](int mt, int r) -> int {  // output pixel index of tile row r, -1 if outside
      if (mode_is_patchy(MODE)) {
        const int y = ((mt / pw_tiles) % ph_tiles) * PH + r / PW, xq = (mt % pw_tiles) * PW + r % PW;
        const int img = mt / (pw_tiles * ph_tiles);
        return (y < ho && xq < wo && img < p.n) ? (img * ho + y) * wo + xq : -1;
      }
      const long long m = (long long)mt * BLOCK_M + r;
      return m < M ? (int)m : -1;
    };
    // The residual tile (128 rows x BN bf16) is pulled into L2 one whole tile ahead with prefetch.global.L2, so the
    // register loads below (issued one sub-tile ahead) hit L2 instead of paying DRAM latency with little in flight.
    constexpr int LINES_PER_ROW = (BN * 2 + 127) / 128;
    auto prefetch_residual = [&](int tile_) {
      if (!has_res || tile_ >= num_tiles) return;
      const int n0_ = (tile_ % num_n_tiles) * BN, mt_ = tile_mt(tile_);
      for (int e = tid; e < BLOCK_M * LINES_PER_ROW; e += EPI_WARPS * 32) {
        const int m = row_to_m(mt_, e / LINES_PER_ROW);
        const int col = n0_ + (e % LINES_PER_ROW) * 64;
        if (m >= 0 && col < p.cout) {
          const char* a = reinterpret_cast<const char*>(p.residual) + ((size_t)m * p.res_ld + col) * esz;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
          if (SPLIT) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + p.res_plane * 2));
        }
      }
    };
    prefetch_residual(my_first);
    const uint32_t empty_rank0 = CTA2 ? map_to_cta(tmem_empty_bar(0), 0) : 0u;      // the leader's MMA thread waits for both epilogues
    int it = ALT ? half : 0;
    for (int tile = my_first; tile < num_tiles; tile += my_step, it += CHUNKED ? 0 : (ALT ? 2 : 1)) {    // (CHUNKED: `it` counts K chunks)
      const int acc = it & 1;
      const Unit u = decode_unit<ACC>(tile, num_m_tiles, num_n_tiles, num_splits, num_kb);
      const int n0 = u.nt * BN;
      const int mt = CTA2 ? 2 * u.mt + cta_rank : u.mt;
      prefetch_residual(tile + my_step);
      int mrow[4];
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) mrow[ps] = row_to_m(mt, quarter * 32 + ps * 8 + rsub);
      // (split mode: the residual is a pair -- dst = hi plane, dst2 = lo plane)
      auto load_res = [&](int cc, uint4 (&dst)[4], uint4 (&dst2)[SPLIT ? 4 : 1]) {
        const int co = n0 + cc * SUB + colv;
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) {
          dst[ps] = make_uint4(0u, 0u, 0u, 0u);
          if (SPLIT) dst2[SPLIT ? ps : 0] = make_uint4(0u, 0u, 0u, 0u);
          if (has_res && aligned && cc < NSUB && mrow[ps] >= 0 && co + 8 <= p.cout) {
            const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + (size_t)mrow[ps] * p.res_ld + co;
            dst[ps] = __ldg(reinterpret_cast<const uint4*>(rp));
            if (SPLIT) dst2[SPLIT ? ps : 0] = __ldg(reinterpret_cast<const uint4*>(rp + p.res_plane));
          }
        }
      };
      // residual of the first sub-tile is requested before the accumulator is even ready
      uint4 rv[4], rv2[SPLIT ? 4 : 1];
      if (!CHUNKED) load_res(sub0, rv, rv2);       // (CHUNKED: after the K-chunk drain, whose loop wants the registers)
      uint32_t v[32];
      // split mode: this warp's (up to two) 32-column sub-tiles accumulate in registers over the K chunks of the tile
      uint32_t racc0[CHUNKED ? 32 : 1], racc1[CHUNKED ? 32 : 1];
      if constexpr (CHUNKED) {
        constexpr int CH = chunk_kb(MODE);
        const int nchunks = (num_kb + CH - 1) / CH;
        for (int c = 0; c < nchunks; ++c, ++it) {
          const int accb = it % NBUF;
          mbar_wait(tmem_full_bar(accb), (it / NBUF) & 1);
          tc_fence_after();
          const uint32_t t_rowc = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(accb * BN);
          // one sub-tile at a time through the same 32 staging registers (both at once would not fit next to the 64 sums)
          if (half < NSUB) tmem_ld32_nowait(t_rowc + (uint32_t)(half * SUB), v);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e) racc0[e] = c == 0 ? v[e] : __float_as_uint(__fadd_rn(__uint_as_float(racc0[e]), __uint_as_float(v[e])));
          if (BN > 64) {
            if (half + 2 < NSUB) tmem_ld32_nowait(t_rowc + (uint32_t)((half + 2) * SUB), v);
            tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 32; ++e) racc1[e] = c == 0 ? v[e] : __float_as_uint(__fadd_rn(__uint_as_float(racc1[e]), __uint_as_float(v[e])));
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (CTA2) mbar_arrive_cluster(empty_rank0 + 8u * accb); else mbar_arrive(tmem_empty_bar(accb)); }
        }
        load_res(half, rv, rv2);
      } else {
      mbar_wait(tmem_full_bar(acc), (it >> 1) & 1);
      tc_fence_after();
      }
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
      if (!CHUNKED && sub0 < NSUB) tmem_ld32_nowait(t_row + (uint32_t)(sub0 * SUB), v);
#pragma unroll 1
      for (int k = 0; k < MY_SUBS; ++k) {
        const int cc = sub0 + 2 * k;
        if (cc >= NSUB) break;
        // bf16: the next sub-tile's residual is requested one sub-tile ahead; split mode (twice the registers per residual)
        // requests it right after this sub-tile's has been consumed
        uint4 rn[4], rn2[1];
        if (!SPLIT) load_res(cc + 2, rn, *reinterpret_cast<uint4(*)[SPLIT ? 4 : 1]>(&rn2));
        // CoordConv bias map rows of this sub-tile: all four passes' loads are issued here, under the TMEM wait and the slab
        // write, instead of one exposed L2 round trip per pass (the map never fits the L1 left beside the operand stages)
        float4 bmv[4][2];
        if (p.bias_map) {
          const int co_b = n0 + cc * SUB + colv;
#pragma unroll
          for (int ps = 0; ps < 4; ++ps) {
            bmv[ps][0] = bmv[ps][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (aligned && mrow[ps] >= 0 && co_b + 8 <= p.cout) {
              const float4* bm = reinterpret_cast<const float4*>(p.bias_map + (size_t)((unsigned)mrow[ps] % hw_out) * p.cout + co_b);
              bmv[ps][0] = __ldg(bm); bmv[ps][1] = __ldg(bm + 1);
            }
          }
        }
        // phase 1: accumulator registers (lane = row) -> swizzled warp-private slab
        if (!CHUNKED) tmem_wait_ld();
        {
          const uint32_t st_row = slab_u32 + (uint32_t)lane * (ST_LD * 4);
          const uint32_t sw = (uint32_t)(lane & 7);
          if (CHUNKED && k == 1) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_row + ((((uint32_t)q) ^ sw) << 4)),
                           "r"(racc1[CHUNKED ? 4 * q : 0]), "r"(racc1[CHUNKED ? 4 * q + 1 : 0]), "r"(racc1[CHUNKED ? 4 * q + 2 : 0]), "r"(racc1[CHUNKED ? 4 * q + 3 : 0]) : "memory");
          } else if (CHUNKED) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_row + ((((uint32_t)q) ^ sw) << 4)),
                           "r"(racc0[CHUNKED ? 4 * q : 0]), "r"(racc0[CHUNKED ? 4 * q + 1 : 0]), "r"(racc0[CHUNKED ? 4 * q + 2 : 0]), "r"(racc0[CHUNKED ? 4 * q + 3 : 0]) : "memory");
          } else {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_row + ((((uint32_t)q) ^ sw) << 4)),
                         "r"(v[4 * q]), "r"(v[4 * q + 1]), "r"(v[4 * q + 2]), "r"(v[4 * q + 3]) : "memory");
          }
        }
        if (CHUNKED) {
          __syncwarp();                              // (the accumulator buffers were released chunk by chunk)
        } else if (cc + 2 < NSUB) {
          tmem_ld32_nowait(t_row + (uint32_t)((cc + 2) * SUB), v);     // next sub-tile streams in during phase 2
          __syncwarp();
        } else {                                   // this warp's TMEM reads of the tile are done: release the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (CTA2) mbar_arrive_cluster(empty_rank0 + 8u * acc); else mbar_arrive(tmem_empty_bar(acc)); }
        }
        // phase 2: coalesced (lane = 8 channels of one row; 8 rows per pass, 4 passes)
        const int co = n0 + cc * SUB + colv;
        if (co < p.cout) {
          if (aligned && co + 8 <= p.cout) {
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.scale + co)), s1 = __ldg(reinterpret_cast<const float4*>(p.scale + co) + 1);
            const float4 h0 = __ldg(reinterpret_cast<const float4*>(p.shift + co)), h1 = __ldg(reinterpret_cast<const float4*>(p.shift + co) + 1);
#pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
              const int m = mrow[ps];
              if (m < 0) continue;
              const int row = ps * 8 + rsub;
              const float* sp = slab + row * ST_LD;
              float4 a = *reinterpret_cast<const float4*>(sp + ((cpair ^ (row & 7)) << 2));
              float4 b = *reinterpret_cast<const float4*>(sp + (((cpair + 1) ^ (row & 7)) << 2));
              if (p.bias_map) {
                const float4 b0 = bmv[ps][0], b1 = bmv[ps][1];
                a.x += b0.x; a.y += b0.y; a.z += b0.z; a.w += b0.w; b.x += b1.x; b.y += b1.y; b.z += b1.z; b.w += b1.w;
              }
              a.x = a.x * s0.x + h0.x; a.y = a.y * s0.y + h0.y; a.z = a.z * s0.z + h0.z; a.w = a.w * s0.w + h0.w;
              b.x = b.x * s1.x + h1.x; b.y = b.y * s1.y + h1.y; b.z = b.z * s1.z + h1.z; b.w = b.w * s1.w + h1.w;
              const uint4 t = rv[ps];            // zeros when there is no residual
              if (SPLIT) {
                if (has_res) add_pair8(a, b, t, rv2[SPLIT ? ps : 0]);
              } else {
              a.x += bf_lo(t.x); a.y += bf_hi(t.x); a.z += bf_lo(t.y); a.w += bf_hi(t.y);
              b.x += bf_lo(t.z); b.y += bf_hi(t.z); b.z += bf_lo(t.w); b.w += bf_hi(t.w);
              }
              if (slope != 1.f) {                // relu / leaky(0.1): max(v, slope*v) is exact for 0 <= slope < 1
                a.x = fmaxf(a.x, a.x * slope); a.y = fmaxf(a.y, a.y * slope); a.z = fmaxf(a.z, a.z * slope); a.w = fmaxf(a.w, a.w * slope);
                b.x = fmaxf(b.x, b.x * slope); b.y = fmaxf(b.y, b.y * slope); b.z = fmaxf(b.z, b.z * slope); b.w = fmaxf(b.w, b.w * slope);
              }
              if (p.upsample2x) {
                store_upsampled(p.y, p.y_ld, p.y_plane, p.out_dtype, p.overflow, m, co, ho, wo, a, b);
              } else if (SPLIT && out_bf16) {
                uint4 hi, lo;
                split8(a, b, hi, lo, p.overflow);
                __half* d = reinterpret_cast<__half*>(p.y) + (size_t)m * p.y_ld + co;
                *reinterpret_cast<uint4*>(d) = hi;
                *reinterpret_cast<uint4*>(d + p.y_plane) = lo;
              } else if (out_bf16) {
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.y) + (size_t)m * p.y_ld + co) =
                    make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
              } else {
                float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + (size_t)m * p.y_ld + co);
                dst[0] = a; dst[1] = b;
              }
            }
          } else {
#pragma unroll 1
            for (int ps = 0; ps < 4; ++ps) {
              if (mrow[ps] < 0) continue;
              const int row = ps * 8 + rsub;
              const float4 a = *reinterpret_cast<const float4*>(slab + row * ST_LD + ((cpair ^ (row & 7)) << 2));
              const float4 b = *reinterpret_cast<const float4*>(slab + row * ST_LD + (((cpair + 1) ^ (row & 7)) << 2));
              if constexpr (ACC) {
                epilogue_acc(reinterpret_cast<float*>(p.y), p.y_ld, p.cout, p.scale, p.shift, a, b, mrow[ps], co, u.tap * p.wgrad_tap_stride, u.sp == 0);
              } else {
                epilogue_slow(*p_smem, a, b, mrow[ps], co, ho, wo, slope);
              }
            }
          }
        }
        if (SPLIT) load_res(cc + 2, rv, rv2);
        else {
#pragma unroll
          for (int ps = 0; ps < 4; ++ps) rv[ps] = rn[ps];
        }
        __syncwarp();                            // slab is rewritten by the next sub-tile
      }
      if (!CHUNKED && !ALT && half >= NSUB) {    // BN == 32 with partial sums: the second warp of a quarter has no sub-tile, still releases
        tc_fence_before();
        if (lane == 0) { if (CTA2) mbar_arrive_cluster(empty_rank0 + 8u * acc); else mbar_arrive(tmem_empty_bar(acc)); }
      }
    }
  }

  tc_fence_before();
  if (CTA2) cluster_sync_all(); else __syncthreads();       // pair: nobody leaves while the peer may still touch its smem / TMEM
  if (warp == MMA_WARP) {
    if (CTA2) tmem_dealloc2(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// Debug knobs (read once): PPY_NO_PDL / PPY_NO_CTA2 / PPY_NO_TMA_EPI / PPY_NO_PATCH / PPY_NO_SLAB / PPY_NO_IM2COL switch one
// mechanism off so tools/conv_bench.py can A/B it on the same GPU; every combination is a correct (slower) kernel.
bool knob_off(const char* name) { return getenv(name) != nullptr; }

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

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeIm2colFn get_encode_im2col_fn() {
  static EncodeIm2colFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeIm2colFn>(ptr);
  }
  return fn;
}

int num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

constexpr CUtensorMapDataType OPERAND_DT = SPLIT ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
constexpr CUtensorMapSwizzle OPERAND_SWIZZLE = K32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;

int encode_2d(EncodeTiledFn enc, CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t row_bytes,
              uint32_t box_inner, uint32_t box_outer, CUtensorMapDataType dt = OPERAND_DT) {
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  const cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  CUresult cr = enc(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, OPERAND_SWIZZLE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { g_last_cuda_error = (int)cr; return PPY_ERR_CUDA; }
  return PPY_OK;
}

int encode_patch_4d(EncodeTiledFn enc, CUtensorMap* map, const ppy_conv_params* p, const void* base, int box_w, int box_h) {
  // NHWC activation as (C, W, H, N); box = 64 channels x box_w x box_h pixels of one image
  const cuuint64_t dims[4] = {(cuuint64_t)p->cin, (cuuint64_t)p->w, (cuuint64_t)p->h, (cuuint64_t)p->n};
  const cuuint64_t strides[3] = {(cuuint64_t)p->x_ld * 2, (cuuint64_t)p->w * p->x_ld * 2, (cuuint64_t)p->h * p->w * p->x_ld * 2};
  const cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult cr = enc(map, OPERAND_DT, 4, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, OPERAND_SWIZZLE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { g_last_cuda_error = (int)cr; return PPY_ERR_CUDA; }
  return PPY_OK;
}

int encode_im2col_4d(CUtensorMap* map, const ppy_conv_params* p, const void* base) {
  // NHWC activation as (C, W, H, N) in im2col mode: a box is 64 channels x BLOCK_M pixels walked from a base pixel with the
  // conv stride; the base pixel's bounding box is [-pad, dim - 1 + pad - (k - 1)] so the walk wraps exactly at wo / ho
  EncodeIm2colFn enc = get_encode_im2col_fn();
  if (!enc) return PPY_ERR_UNSUPPORTED;
  const cuuint64_t dims[4] = {(cuuint64_t)p->cin, (cuuint64_t)p->w, (cuuint64_t)p->h, (cuuint64_t)p->n};
  const cuuint64_t strides[3] = {(cuuint64_t)p->x_ld * 2, (cuuint64_t)p->w * p->x_ld * 2, (cuuint64_t)p->h * p->w * p->x_ld * 2};
  const int lower[2] = {-p->pad, -p->pad};
  const int upper[2] = {p->pad - (p->kw - 1), p->pad - (p->kh - 1)};
  const cuuint32_t estr[4] = {1, (cuuint32_t)p->stride, (cuuint32_t)p->stride, 1};
  CUresult cr = enc(map, OPERAND_DT, 4, const_cast<void*>(base), dims, strides, lower, upper, (cuuint32_t)BLOCK_K,
                    (cuuint32_t)BLOCK_M, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, OPERAND_SWIZZLE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { g_last_cuda_error = (int)cr; return PPY_ERR_CUDA; }
  // drivers up to CUDA 13.1 set a descriptor bit for tensors under 128 KB that im2col loads then mishandle (same fix-up as CUTLASS)
  int drv = 0;
  if (cudaDriverGetVersion(&drv) == cudaSuccess && drv <= 13010 && (long long)p->n * p->h * p->w * p->x_ld * 2 < 131072)
    reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);
  return PPY_OK;
}

int encode_tile_map(EncodeTiledFn enc, CUtensorMap* map, const void* base, int ld, int cols, const ppy_conv_params* p, int ho,
                    int wo, bool patch, int box_w, int box_h, bool f32 = false) {
  // output / residual tensors: channels innermost; 2-D [M rows][cols] or 4-D (C, W, H, N) for patch tiles.  A box row is always
  // 128 bytes: 64 bf16 channels, or 32 fp32 channels (fp32 outputs leave as two boxes per 64-column group)
  const uint64_t esz = f32 ? 4 : 2;
  const uint32_t box_cols = f32 ? GROUP_COLS / 2 : GROUP_COLS;
  const CUtensorMapDataType dt = f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : OPERAND_DT;
  if (!patch) return encode_2d(enc, map, base, (uint64_t)cols, (uint64_t)p->n * ho * wo, (uint64_t)ld * esz, box_cols, BOX_ROWS, dt);
  const cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)wo, (cuuint64_t)ho, (cuuint64_t)p->n};
  const cuuint64_t strides[3] = {(cuuint64_t)ld * esz, (cuuint64_t)wo * ld * esz, (cuuint64_t)ho * wo * ld * esz};
  const cuuint32_t box[4] = {box_cols, (cuuint32_t)box_w, (cuuint32_t)(BOX_ROWS / box_w), 1};   // one warp's 32 tile rows
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult cr = enc(map, dt, 4, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, OPERAND_SWIZZLE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { g_last_cuda_error = (int)cr; return PPY_ERR_CUDA; }
  return PPY_OK;
}

// K splits for accumulate launches (fp32 atomics into a zeroed output): minimise rounds x (k-blocks per unit + fixed
// per-unit cost of the pipeline ramp and epilogue, ~6 k-blocks) over the split counts that leave no unit empty.
int pick_splits(const ppy_conv_params* p, long long tiles, int num_kb) {
  if (!p->accumulate) return 1;
  auto normalise = [&](int want) {
    if (want < 1) want = 1;
    if (want > num_kb) want = num_kb;
    const int per = (num_kb + want - 1) / want;
    return (num_kb + per - 1) / per;
  };
  if (p->split_k > 0) return normalise(p->split_k);
  int best = 1;
  long long best_cost = -1;
  for (int s = 1; s <= 32 && s <= num_kb; ++s) {
    const int sn = normalise(s);
    const long long rounds = ceil_div(tiles * sn, num_sms());
    const long long cost = rounds * (ceil_div(num_kb, sn) + 6);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = sn; }
  }
  return best;
}

template <int BN, int MODE, int EPI, bool ACC = false, bool CTA2 = false, bool CHUNKED = false>
int launch(const ppy_conv_params* p, int ho, int wo, cudaStream_t st) {
  using Cfg = TileCfg<BN, MODE, EPI, CTA2>;
  constexpr int PW = patch_w(MODE), PH = patch_h(MODE);
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return PPY_ERR_UNSUPPORTED;
  CUtensorMap tmap_b, tmap_a, tmap_y, tmap_r, tmap_a2;
  // split mode: the packed weight is [2][cout_pad][k_pad] (hi plane, lo plane): one map, the lo rows start at cout_pad
  int rc = encode_2d(enc, &tmap_b, p->weight, (uint64_t)p->k_pad, (uint64_t)PLANES * p->cout_pad, (uint64_t)p->k_pad * 2, BLOCK_K, CTA2 ? BN / 2 : BN);
  if (rc) return rc;
  const long long M = (long long)p->n * ho * wo;
  tmap_a = tmap_b;
  tmap_a2 = tmap_b;
  for (int pl = 0; pl < PLANES; ++pl) {          // split mode: a second map over the lo plane of x
    CUtensorMap* ma = pl ? &tmap_a2 : &tmap_a;
    const void* xb = reinterpret_cast<const uint16_t*>(p->x) + (pl ? p->x_plane : 0);
    if (MODE == MODE_TMA_A) {
      // 1x1 stride-1: the A operand is the NHWC activation itself, [M rows][cin] with row pitch x_ld
      rc = encode_2d(enc, ma, xb, (uint64_t)p->cin, (uint64_t)M, (uint64_t)p->x_ld * 2, BLOCK_K, BLOCK_M);
    } else if (MODE == MODE_TMA_PATCH) {
      rc = encode_patch_4d(enc, ma, p, xb, PW, PH);
    } else if (MODE == MODE_TMA_SLAB) {
      rc = encode_patch_4d(enc, ma, p, xb, PW, SLAB_ROWS);
    } else if (MODE == MODE_TMA_IM2COL) {
      rc = encode_im2col_4d(ma, p, xb);
    }
    if (rc) return rc;
  }
  tmap_y = tmap_b;
  tmap_r = tmap_b;
  if (SPLIT && p->x2_kb > 0) {      // second K source: hi plane -> tmap_y, lo plane -> tmap_r
    if (!(MODE == MODE_TMA_A || MODE == MODE_TMA_IM2COL || MODE == MODE_TMA_PATCH)) return PPY_ERR_UNSUPPORTED;
    const uint64_t c2 = p->x2_tiled ? (uint64_t)p->cout : (uint64_t)p->x2_kb * BLOCK_K;
    // x2_row_mod > 0: a batch-invariant source of x2_row_mod rows per image, stored x2_rows >= x2_row_mod + 127 rows long
    const uint64_t rows2 = p->x2_row_mod ? (uint64_t)p->x2_rows : (uint64_t)M;
    for (int pl = 0; pl < 2; ++pl) {
      const void* b2 = reinterpret_cast<const uint16_t*>(p->x2) + (pl ? p->x2_plane : 0);
      CUtensorMap* m2 = pl ? &tmap_r : &tmap_y;
      if (MODE == MODE_TMA_PATCH) {
        const cuuint64_t dims[4] = {(cuuint64_t)c2, (cuuint64_t)wo, (cuuint64_t)ho, (cuuint64_t)(p->x2_row_mod ? 1 : p->n)};
        const cuuint64_t strides[3] = {(cuuint64_t)p->x2_ld * 2, (cuuint64_t)wo * p->x2_ld * 2, (cuuint64_t)ho * wo * p->x2_ld * 2};
        const cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)PW, (cuuint32_t)PH, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult cr = enc(m2, OPERAND_DT, 4, const_cast<void*>(b2), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          OPERAND_SWIZZLE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) { g_last_cuda_error = (int)cr; return PPY_ERR_CUDA; }
      } else {
        rc = encode_2d(enc, m2, b2, c2, rows2, (uint64_t)p->x2_ld * 2, BLOCK_K, BLOCK_M);
        if (rc) return rc;
      }
    }
  }
  if (EPI == EPI_TMA) {
    rc = encode_tile_map(enc, &tmap_y, p->y, p->y_ld, p->cout, p, ho, wo, mode_is_patchy(MODE), PW, PH, p->out_dtype == PPY_F32);
    if (rc) return rc;
    if (p->residual) {
      rc = encode_tile_map(enc, &tmap_r, p->residual, p->res_ld, p->cout, p, ho, wo, mode_is_patchy(MODE), PW, PH);
      if (rc) return rc;
    }
  }
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    rc = check_cuda(cudaFuncSetAttribute(conv_umma_kernel<BN, MODE, EPI, ACC, CTA2, CHUNKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    if (rc) return rc;
  }
  const int pw_tiles = (int)ceil_div(wo, PW), ph_tiles = (int)ceil_div(ho, PH);
  int num_m_tiles = mode_is_patchy(MODE) ? p->n * pw_tiles * ph_tiles : (int)ceil_div(M, BLOCK_M);
  if (CTA2) num_m_tiles = (num_m_tiles + 1) / 2;            // scheduler units are pairs of M tiles
  const int num_n_tiles = (int)ceil_div(p->cout, BN);
  const int num_kb = MODE == MODE_TMA_SLAB ? 3 * (p->cin / BLOCK_K) : p->k_pad / BLOCK_K;   // SLAB: one iteration per (channel block, kx)
  const int num_taps = p->wgrad_taps > 0 ? p->wgrad_taps : 1;
  const int num_splits = pick_splits(p, (long long)num_m_tiles * num_n_tiles * num_taps, num_kb);
  const long long tiles = (long long)num_m_tiles * num_n_tiles * num_taps * num_splits;
  static const bool use_pdl = !knob_off("PPY_NO_PDL");
  const int pairs = num_sms() / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = CTA2 ? dim3((unsigned)(2 * (tiles < pairs ? tiles : pairs))) : dim3((unsigned)(tiles < num_sms() ? tiles : num_sms()));
  cfg.blockDim = dim3(num_threads(MODE));
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CTA2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (use_pdl) {                     // prologue overlaps the previous kernel's tail (griddepcontrol.wait in the kernel)
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  rc = check_cuda(cudaLaunchKernelEx(&cfg, conv_umma_kernel<BN, MODE, EPI, ACC, CTA2, CHUNKED>, *p, ho, wo, num_kb, num_m_tiles, num_n_tiles,
                                     num_splits, num_taps, pw_tiles, ph_tiles, tmap_b, tmap_a, tmap_y, tmap_r, tmap_a2));
  if (rc) return rc;
  return check_launch();
}

// The TMA epilogue needs bf16 output, 16-byte aligned rows, no CoordConv bias map / fused upsample, and at least one
// 64-column group; it is used for K <= 512 (the HBM-bound layers), the slab epilogue with its deeper operand ring elsewhere.
bool tma_epilogue_ok(const ppy_conv_params* p) {
  static const bool no_tma_epi = knob_off("PPY_NO_TMA_EPI");
  if (p->accumulate || no_tma_epi || SPLIT) return false;      // (split mode has the slab epilogue only)
  if (p->bias_map || p->upsample2x || p->cout < GROUP_COLS) return false;
  if (p->out_dtype == PPY_F32 && p->residual) return false;              // fp32 outputs (head output convs): two boxes per group, no residual
  if (p->k_pad > ((p->cout % 256 == 0) ? 512 : 1152)) return false;     // 3 operand stages at BLOCK_N 256, 4-6 below
  if ((reinterpret_cast<uintptr_t>(p->y) & 15) || (p->y_ld * dtype_size(p->out_dtype)) % 16) return false;
  if (p->residual && ((reinterpret_cast<uintptr_t>(p->residual) & 15) || (p->res_ld * 2) % 16)) return false;
  return true;
}

template <int MODE>
int dispatch(const ppy_conv_params* p, int ho, int wo, cudaStream_t st) {
  const int c = p->cout;
  if constexpr (SPLIT) {     // fp16-pair operands: slab epilogue only, no partial sums
    if (p->accumulate) return PPY_ERR_UNSUPPORTED;
    static const bool no_pair = knob_off("PPY_NO_CTA2");
    const bool pair = mode_is_tma(MODE) && !no_pair && (long long)p->n * ho * wo > BLOCK_M;
    if constexpr (MODE == MODE_TMA_A) {
      if (p->x2_kb > 0 && p->x2_tiled) {     // identity blocks: the N tile is fixed by their width; up to 8 K blocks in one accumulator
        const int kbt = p->k_pad / BLOCK_K;
        if (p->x2_kb == 4 && c % 256 == 0 && kbt <= 8)
          return pair ? launch<256, MODE, EPI_SLAB, false, true>(p, ho, wo, st) : launch<256, MODE, EPI_SLAB>(p, ho, wo, st);
        if (p->x2_kb == 2 && c % 128 == 0) {
          if (kbt <= 8) return pair ? launch<128, MODE, EPI_SLAB, false, true>(p, ho, wo, st) : launch<128, MODE, EPI_SLAB>(p, ho, wo, st);
          return pair ? launch<128, MODE, EPI_SLAB, false, true, true>(p, ho, wo, st) : launch<128, MODE, EPI_SLAB, false, false, true>(p, ho, wo, st);
        }
        return PPY_ERR_UNSUPPORTED;
      }
    }
    if (p->k_pad / BLOCK_K <= (p->x2_kb > 0 ? 6 : chunk_kb(MODE))) {    // the whole K is one chunk (K <= 256): no register stage, BLOCK_N up to 256
      if (c <= 32) return launch<32, MODE, EPI_SLAB>(p, ho, wo, st);
      if (c <= 64) return launch<64, MODE, EPI_SLAB>(p, ho, wo, st);
      if constexpr (mode_is_tma(MODE)) {
        if (pair) return c % 256 == 0 ? launch<256, MODE, EPI_SLAB, false, true>(p, ho, wo, st) : launch<128, MODE, EPI_SLAB, false, true>(p, ho, wo, st);
      }
      return c % 256 == 0 ? launch<256, MODE, EPI_SLAB>(p, ho, wo, st) : launch<128, MODE, EPI_SLAB>(p, ho, wo, st);
    }
    // K chunks summed in the epilogue warps' registers: BLOCK_N <= 128
    if (c <= 32) return launch<32, MODE, EPI_SLAB, false, false, true>(p, ho, wo, st);
    if (c <= 64) return launch<64, MODE, EPI_SLAB, false, false, true>(p, ho, wo, st);
    if constexpr (mode_is_tma(MODE)) {
      if (pair) return launch<128, MODE, EPI_SLAB, false, true, true>(p, ho, wo, st);
    }
    return launch<128, MODE, EPI_SLAB, false, false, true>(p, ho, wo, st);
  } else {
  if (p->accumulate) {       // partial-sum launches (split-K, weight gradients): slab epilogue with fp32 atomics
    if (MODE == MODE_DCN || mode_is_patchy(MODE)) return PPY_ERR_UNSUPPORTED;
    constexpr int M2 = (MODE == MODE_TMA_A || MODE == MODE_TMA_IM2COL) ? MODE : MODE_GATHER;
    if (c <= 32) return launch<32, M2, EPI_SLAB, true>(p, ho, wo, st);
    if (c <= 64) return launch<64, M2, EPI_SLAB, true>(p, ho, wo, st);
    if (c <= 128) return launch<128, M2, EPI_SLAB, true>(p, ho, wo, st);
    return launch<256, M2, EPI_SLAB, true>(p, ho, wo, st);
  }
  if (c <= 32) return launch<32, MODE, EPI_SLAB>(p, ho, wo, st);
  const bool tma_epi = MODE != MODE_DCN && tma_epilogue_ok(p);
  if (c <= 64) return tma_epi ? launch<64, MODE, EPI_TMA>(p, ho, wo, st) : launch<64, MODE, EPI_SLAB>(p, ho, wo, st);
  if constexpr (mode_is_tma(MODE)) {
    // CTA pairs (cta_group::2, 256 x BN tiles, B split across the pair) for every TMA-fed layer with at least two M tiles
    static const bool no_pair = knob_off("PPY_NO_CTA2");
    if (!no_pair && (long long)p->n * ho * wo > BLOCK_M) {
      // few tiles (small batches: the 19 x 19 maps of a bs-8 training step give 24 pair tiles at BLOCK_N 256 -- 48 of 148 SMs):
      // 128-wide tiles when twice as many CTAs still fit one wave
      static const bool no_narrow = knob_off("PPY_NO_NARROW_TILES");
      const long long ctas256 = 2 * ceil_div((long long)p->n * ho * wo, 2ll * BLOCK_M) * (c / 256);
      const bool narrow = !no_narrow && c % 256 == 0 && 2 * ctas256 <= num_sms();
      if (c % 256 == 0 && !narrow)
        return tma_epi ? launch<256, MODE, EPI_TMA, false, true>(p, ho, wo, st) : launch<256, MODE, EPI_SLAB, false, true>(p, ho, wo, st);
      return tma_epi ? launch<128, MODE, EPI_TMA, false, true>(p, ho, wo, st) : launch<128, MODE, EPI_SLAB, false, true>(p, ho, wo, st);
    }
  }
  if (c % 256 == 0) return tma_epi ? launch<256, MODE, EPI_TMA>(p, ho, wo, st) : launch<256, MODE, EPI_SLAB>(p, ho, wo, st);
  return tma_epi ? launch<128, MODE, EPI_TMA>(p, ho, wo, st) : launch<128, MODE, EPI_SLAB>(p, ho, wo, st);
  }
}

// 3x3 stride-1 convs with 64..128 output channels (the stem and the stage-2/3 bottleneck 3x3s): slab stages, CTA pairs
template <int UNUSED = 0>     // (a template so that `if constexpr` discards the branches of the other builds)
int dispatch_slab(const ppy_conv_params* p, int ho, int wo, cudaStream_t st) {
  static const bool no_pair = knob_off("PPY_NO_CTA2");
  const bool tma_epi = tma_epilogue_ok(p);
  const bool pair = !no_pair && p->n * ceil_div(ho, patch_h(MODE_TMA_SLAB)) * ceil_div(wo, patch_w(MODE_TMA_SLAB)) > 1;
  if constexpr (SPLIT && K32) {     // cin = 32: three slab stages (54 MMAs) = one accumulator, no register stage
    if (p->cin != BLOCK_K) return PPY_ERR_UNSUPPORTED;
    if (p->cout <= 32) return launch<32, MODE_TMA_SLAB, EPI_SLAB>(p, ho, wo, st);
    if (p->cout <= 64) return pair ? launch<64, MODE_TMA_SLAB, EPI_SLAB, false, true>(p, ho, wo, st) : launch<64, MODE_TMA_SLAB, EPI_SLAB>(p, ho, wo, st);
    return pair ? launch<128, MODE_TMA_SLAB, EPI_SLAB, false, true>(p, ho, wo, st) : launch<128, MODE_TMA_SLAB, EPI_SLAB>(p, ho, wo, st);
  } else if constexpr (SPLIT) {     // (3x3: always more than one K chunk)
    if (p->cout <= 64) return pair ? launch<64, MODE_TMA_SLAB, EPI_SLAB, false, true, true>(p, ho, wo, st) : launch<64, MODE_TMA_SLAB, EPI_SLAB, false, false, true>(p, ho, wo, st);
    if (pair) return launch<128, MODE_TMA_SLAB, EPI_SLAB, false, true, true>(p, ho, wo, st);
    return dispatch<MODE_TMA_IM2COL>(p, ho, wo, st);            // single-tile problem: a 128-wide slab stage pair does not fit twice
  } else {
  if (p->cout <= 64) {
    if (pair) return tma_epi ? launch<64, MODE_TMA_SLAB, EPI_TMA, false, true>(p, ho, wo, st) : launch<64, MODE_TMA_SLAB, EPI_SLAB, false, true>(p, ho, wo, st);
    return tma_epi ? launch<64, MODE_TMA_SLAB, EPI_TMA>(p, ho, wo, st) : launch<64, MODE_TMA_SLAB, EPI_SLAB>(p, ho, wo, st);
  }
  if (pair) return tma_epi ? launch<128, MODE_TMA_SLAB, EPI_TMA, false, true>(p, ho, wo, st) : launch<128, MODE_TMA_SLAB, EPI_SLAB, false, true>(p, ho, wo, st);
  return launch<128, MODE_TMA_SLAB, EPI_SLAB>(p, ho, wo, st);     // single-tile problems only: no room for the TMA epilogue's boxes
  }
}

}  // namespace
}  // namespace ppy

extern "C" {
using namespace ppy;

#if PPY_UMMA_K32
int ppy_conv_bf16_supported(void);

// 3x3 stride-1 pad-1 convs with 32 input channels on pair operands (called by ppy_conv_f16x2)
int ppy_conv_f16x2_k32(const ppy_conv_params* p, ppy_stream_t s) {
  int ho, wo;
  int rc = validate_conv(p, 2, &ho, &wo);
  if (rc) return rc;
  PPY_REQUIRE(SPLIT && p->kh == 3 && p->stride == 1 && p->pad == 1 && p->cin == BLOCK_K && p->k_pad % 64 == 0 && p->k_pad >= 9 * BLOCK_K);
  PPY_REQUIRE(p->out_dtype == PPY_F16X2 || p->out_dtype == PPY_F32);
  PPY_REQUIRE(p->x_plane > 0 && (p->x_plane * 2) % 16 == 0 && !p->accumulate && !p->coord_w && !p->offset_mask && p->x2_kb == 0);
  if (p->out_dtype == PPY_F16X2) PPY_REQUIRE(p->y_plane > 0);
  if (p->residual) PPY_REQUIRE(p->out_dtype == PPY_F16X2 && p->res_plane > 0);
  PPY_REQUIRE((long long)p->n * ho * wo < 0x7FFFFFFFll && p->cout <= 128 && (reinterpret_cast<uintptr_t>(p->x) & 15) == 0);
  PPY_REQUIRE((reinterpret_cast<uintptr_t>(p->scale) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->shift) & 15) == 0);
  if (p->act == PPY_ACT_MISH || !ppy_conv_bf16_supported()) return PPY_ERR_UNSUPPORTED;
  return dispatch_slab(p, ho, wo, as_stream(s));
}
#elif !PPY_UMMA_SPLIT
int ppy_conv_bf16_supported(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 && get_encode_fn() != nullptr;
}

int ppy_conv_bf16(const ppy_conv_params* p, ppy_stream_t s) {
#else
int ppy_conv_bf16_supported(void);
int ppy_conv_f16x2_k32(const ppy_conv_params* p, ppy_stream_t s);

int ppy_conv_f16x2(const ppy_conv_params* p, ppy_stream_t s) {
#endif
#if !PPY_UMMA_K32
  int ho, wo;
  int rc = validate_conv(p, 2, &ho, &wo);
  if (rc) return rc;
  PPY_REQUIRE(p->k_pad * 2 % 16 == 0);
  if (SPLIT) {       // pairs in, pair or fp32 out; the lo planes obey the same alignment as the hi planes
    PPY_REQUIRE(p->out_dtype == PPY_F16X2 || p->out_dtype == PPY_F32);
    PPY_REQUIRE(p->x_plane > 0 && (p->x_plane * 2) % 16 == 0 && !p->accumulate && !p->coord_w);
    if (p->out_dtype == PPY_F16X2) PPY_REQUIRE(p->y_plane > 0);
    if (p->residual) PPY_REQUIRE(p->out_dtype == PPY_F16X2 && p->res_plane > 0);
    if (p->x2_kb > 0) {
      PPY_REQUIRE(p->x2 && !p->residual && !p->offset_mask && p->stride == 1 && p->cin % BLOCK_K == 0);
      PPY_REQUIRE(p->kh == 1 || (!p->x2_tiled && p->kh == 3 && p->pad == 1));
      PPY_REQUIRE(p->k_pad == p->kh * p->kw * p->cin + p->x2_kb * BLOCK_K && p->x2_plane > 0 && (p->x2_plane * 2) % 16 == 0);
      if (p->x2_row_mod) PPY_REQUIRE(p->x2_row_mod == ho * wo && p->x2_rows >= p->x2_row_mod + BLOCK_M - 1 && !p->x2_tiled);
      PPY_REQUIRE((reinterpret_cast<uintptr_t>(p->x2) & 15) == 0 && (p->x2_ld * 2) % 16 == 0);
      PPY_REQUIRE(p->x2_ld >= (p->x2_tiled ? p->cout : p->x2_kb * BLOCK_K));
    }
  } else {
    PPY_REQUIRE(p->out_dtype != PPY_F16X2 && p->x2_kb == 0);
  }
  if (p->offset_mask) PPY_REQUIRE(p->cin % BLOCK_K == 0 && (long long)p->n * p->h * p->w * p->x_ld < 0x7FFFFFFFll);
  PPY_REQUIRE((long long)p->n * ho * wo < 0x7FFFFFFFll);
  if (p->act == PPY_ACT_MISH) return PPY_ERR_UNSUPPORTED;   // no config uses it; ppy_activation covers module-level Mish
  if (p->accumulate) {
    // partial sums (K splits and/or the taps of a weight-gradient GEMM) are added atomically into a caller-zeroed fp32 output
    PPY_REQUIRE(p->out_dtype == PPY_F32 && p->act == PPY_ACT_NONE && !p->residual && !p->bias_map && !p->upsample2x && !p->offset_mask);
    PPY_REQUIRE(p->wgrad_taps == 0 || p->wgrad_taps == 1 || p->wgrad_taps == 9);
    PPY_REQUIRE(p->wgrad_taps <= 1 || (p->kh == 1 && p->stride == 1 && p->wgrad_pitch > 0 && p->wgrad_tap_stride >= p->cout));
  } else {
    PPY_REQUIRE(p->split_k <= 1 && p->wgrad_taps <= 1);
  }
  PPY_REQUIRE((reinterpret_cast<uintptr_t>(p->scale) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->shift) & 15) == 0);
  if (!ppy_conv_bf16_supported()) return PPY_ERR_UNSUPPORTED;
  if (p->offset_mask) {     // whole-layer fused DCNv2 kernel (dcn_umma.cu) when the layer fits it, else the generic producer mode
    const int rc2 = dcn_umma_try(p, ho, wo, SPLIT, as_stream(s));
    if (rc2 != 1) return rc2;
    return dispatch<MODE_DCN>(p, ho, wo, as_stream(s));
  }
#if PPY_UMMA_SPLIT
  if (p->kh == 3 && p->stride == 1 && p->pad == 1 && p->cin == 32 && p->cout <= 128 && p->x2_kb == 0 && !knob_off("PPY_NO_K32")) {
    // 32 input channels: one tap = one 32-element K block in the SWIZZLE_64B build (no structural zeros)
    const double eff = (double)ho * wo / ((double)ceil_div(ho, patch_h(MODE_TMA_SLAB)) * patch_h(MODE_TMA_SLAB) * ceil_div(wo, patch_w(MODE_TMA_SLAB)) * patch_w(MODE_TMA_SLAB));
    if (eff >= 0.85 && (reinterpret_cast<uintptr_t>(p->x) & 15) == 0) return ppy_conv_f16x2_k32(p, s);
  }
#endif
  const bool plain_1x1 = p->kh == 1 && p->stride == 1 && p->pad == 0 && p->cin % BLOCK_K == 0 &&
                         p->k_pad == p->cin + p->x2_kb * BLOCK_K;
  if (plain_1x1) return dispatch<MODE_TMA_A>(p, ho, wo, as_stream(s));
  // 3x3 stride-1: A tiles as 16x8 pixel patches fetched by 4-D TMA, when the patch grid wastes < 15% of the tiles
  const bool patchable = p->kh == 3 && p->stride == 1 && p->pad == 1 && p->cin % BLOCK_K == 0 && p->k_pad == 9 * p->cin + p->x2_kb * BLOCK_K &&
                         (reinterpret_cast<uintptr_t>(p->x) & 15) == 0;
  static const bool no_patch = knob_off("PPY_NO_PATCH"), no_slab = knob_off("PPY_NO_SLAB"), no_im2col = knob_off("PPY_NO_IM2COL");
  if (patchable && !p->accumulate && !no_patch) {
    auto grid_eff = [&](int pw, int ph) { return (double)ho * wo / ((double)ceil_div(ho, ph) * ph * ceil_div(wo, pw) * pw); };
    if (p->cout > 32 && p->cout <= 128 && grid_eff(patch_w(MODE_TMA_SLAB), patch_h(MODE_TMA_SLAB)) >= 0.85 && !no_slab && p->x2_kb == 0)
      return dispatch_slab(p, ho, wo, as_stream(s));
    if (grid_eff(patch_w(MODE_TMA_PATCH), patch_h(MODE_TMA_PATCH)) >= 0.85) return dispatch<MODE_TMA_PATCH>(p, ho, wo, as_stream(s));
  }
  // any other k x k conv over whole 64-channel blocks: im2col-mode TMA (stride and zero padding done by the copy engine)
  const bool im2col_ok = p->cin % BLOCK_K == 0 && p->k_pad == p->kh * p->kw * p->cin + p->x2_kb * BLOCK_K && (reinterpret_cast<uintptr_t>(p->x) & 15) == 0 &&
                         (p->x_ld * 2) % 16 == 0 && p->pad <= 8 && p->kh <= 8 && p->kw <= 8 && p->stride <= 8 &&
                         p->wgrad_taps <= 1 && get_encode_im2col_fn() != nullptr && !no_im2col;
  if (im2col_ok) return dispatch<MODE_TMA_IM2COL>(p, ho, wo, as_stream(s));
  if (p->x2_kb > 0) return PPY_ERR_UNSUPPORTED;
  return dispatch<MODE_GATHER>(p, ho, wo, as_stream(s));
}
#endif

}  // extern "C"
