// fp32 SIMT implicit-GEMM convolution with fused scale/shift (+CoordConv bias map, +residual) and
// activation: the arithmetic-exact (fp32 multiply, fp32 accumulate) path used for the 1e-4 parity
// gate, and the weight packer shared with the tcgen05 path.  No im2col buffer: the A tile is gathered
// straight from the NHWC input (zero fill outside the image), or -- for DCNv2 -- bilinearly sampled
// at the learned offsets and modulated, exactly like the tcgen05 producer does in bf16.
//
//   GEMM view: M = n*ho*wo output pixels, N = cout, K = kh*kw*cin with K index (ky*kw + kx)*cin + c.
//   Tile 64x64x16, 256 threads, 4x4 register block per thread, shared-memory staged.
#include <cuda_fp16.h>
#include "common.cuh"

namespace ppy {
namespace {

constexpr int BM = 64, BN = 64, BK = 16, THREADS = 256;

template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, int cout, int cin_total, int kh, int kw, int c_begin,
                                   int c_count, T* __restrict__ out, int cout_pad, int cin_pad, int k_pad) {
  // one row (output channel) per blockIdx.y, 32-bit index arithmetic inside the row (two 64-bit divisions per element were most
  // of these small launches)
  const int taps = kh * kw;
  for (int co = blockIdx.y; co < cout_pad; co += gridDim.y) {
    T* row = out + (size_t)co * k_pad;
    const float* wrow = w + ((size_t)co * cin_total + c_begin) * taps;
    for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < (unsigned)k_pad; k += gridDim.x * blockDim.x) {
      const unsigned tap = k / (unsigned)cin_pad, c = k - tap * (unsigned)cin_pad;
      float v = 0.f;
      if (co < cout && tap < (unsigned)taps && c < (unsigned)c_count) v = __ldg(wrow + (size_t)c * taps + tap);
      row[k] = from_f<T>(v);
    }
  }
}

// Packing of the INPUT-GRADIENT conv's weight straight from the forward weight (conv_autograd._dgrad): the dgrad conv reads dY
// (o_pad channels) and produces c_main channels with wt[ci][co][ky][kx] = w[co][ci][k-1-ky][k-1-kx], so
// packed[ci][tap * o_pad + co] = w[co][ci][taps - 1 - tap] -- one launch instead of flip + transposing copy + pack.  32 x 32
// (co, ci) tiles through shared memory: reads walk ci (contiguous for 1x1, the taps of a row are one contiguous run), writes walk
// co.  blockIdx.z = tap; the extra z slice zero-fills the K tail [taps * o_pad, k_pad).
__global__ void __launch_bounds__(256) pack_weight_dgrad_kernel(const float* __restrict__ w, int cout, int cin_total, int taps, int c_main,
                                                                __nv_bfloat16* __restrict__ out, int rows_pad, int o_pad, int k_pad) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int tap = blockIdx.z;
  const int ci0 = blockIdx.y * 32, co0 = blockIdx.x * 32;
  if (tap == taps) {                     // K tail
    const int tail = k_pad - taps * o_pad;
    for (int r = ty; r < 32; r += 8) {
      const int ci = ci0 + r;
      for (int col = co0 + tx; col < tail && ci < rows_pad; col += gridDim.x * 32) out[(size_t)ci * k_pad + taps * o_pad + col] = from_f<__nv_bfloat16>(0.f);
    }
    return;
  }
  const int src_tap = taps - 1 - tap;
  for (int r = ty; r < 32; r += 8) {     // r = co within the tile, tx = ci
    const int co = co0 + r, ci = ci0 + tx;
    tile[r][tx] = (co < cout && ci < c_main) ? __ldg(w + ((size_t)co * cin_total + ci) * taps + src_tap) : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {     // r = ci within the tile, tx = co
    const int ci = ci0 + r, co = co0 + tx;
    if (ci < rows_pad && co < o_pad) out[(size_t)ci * k_pad + (size_t)tap * o_pad + co] = from_f<__nv_bfloat16>(tile[tx][r]);
  }
}

// PPY_F16X2 packing: [2][cout_pad][k_pad] fp16 -- hi plane, then lo plane (w ~ hi + lo to 2^-22).  The caller pre-scales every
// output channel by a power of two so that the lo parts stay clear of the fp16 subnormal range.
__global__ void pack_weight_pair_kernel(const float* __restrict__ w, int cout, int cin_total, int kh, int kw, int c_begin,
                                        int c_count, __half* __restrict__ out, int cout_pad, int cin_pad, int k_pad) {
  const long long total = (long long)cout_pad * k_pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i / k_pad), k = (int)(i % k_pad);
    const int tap = k / cin_pad, c = k % cin_pad;
    float v = 0.f;
    if (co < cout && tap < kh * kw && c < c_count)
      v = __ldg(w + (((long long)co * cin_total + c_begin + c) * kh + tap / kw) * kw + tap % kw);
    const __half hi = __float2half_rn(v);
    out[i] = hi;
    out[total + i] = __float2half_rn(v - __half2float(hi));
  }
}

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <bool DCN>
__global__ void __launch_bounds__(THREADS) conv_simt_kernel(ppy_conv_params p, int ho, int wo, int k_true) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const long long M = (long long)p.n * ho * wo;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const float* x = reinterpret_cast<const float*>(p.x);
  const float* wgt = reinterpret_cast<const float*>(p.weight);

  // A loader: row = tid/4, four consecutive k starting at (tid%4)*4
  const int a_row = tid >> 2, a_k = (tid & 3) * 4;
  const long long am = m0 + a_row;
  const bool a_valid = am < M;
  int a_img = 0, a_oy = 0, a_ox = 0;
  if (a_valid) { a_ox = (int)(am % wo); a_oy = (int)((am / wo) % ho); a_img = (int)(am / ((long long)wo * ho)); }
  const float* om_row = DCN && a_valid ? p.offset_mask + am * p.om_ld : nullptr;
  // B loader: cout row = tid/4, four consecutive k
  const int b_row = tid >> 2, b_k = (tid & 3) * 4;
  const bool b_valid = (n0 + b_row) < p.cout_pad;
  const float* b_ptr = wgt + (long long)(n0 + b_row) * p.k_pad + b_k;

  const int ty = tid >> 4, tx = tid & 15;  // 16x16 threads, 4x4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.k_pad; k0 += BK) {
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
    const int k = k0 + a_k;
    if (a_valid && k < k_true) {
      const int tap = k / p.cin, c = k % p.cin;
      const int ky = tap / p.kw, kx = tap % p.kw;
      if (!DCN) {
        const int iy = a_oy * p.stride - p.pad + ky, ix = a_ox * p.stride - p.pad + kx;
        if (iy >= 0 && iy < p.h && ix >= 0 && ix < p.w) av = ld4(x + (((long long)a_img * p.h + iy) * p.w + ix) * p.x_ld + c);
      } else {
        const float dy = __ldg(om_row + 2 * tap), dx = __ldg(om_row + 2 * tap + 1);
        const float ml = __ldg(om_row + 2 * p.kh * p.kw + tap);
        const float mask = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-ml)));
        const float py = (float)(a_oy * p.stride - p.pad + ky) + dy;
        const float px = (float)(a_ox * p.stride - p.pad + kx) + dx;
        const float fy = floorf(py), fx = floorf(px);
        const float ly = py - fy, lx = px - fx, hy = 1.f - ly, hx = 1.f - lx;
        const int y0 = (int)fy, x0 = (int)fx;
        const float wgt4[4] = {hy * hx, hy * lx, ly * hx, ly * lx};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int yy = y0 + (q >> 1), xx = x0 + (q & 1);
          if (yy >= 0 && yy < p.h && xx >= 0 && xx < p.w) {
            const float4 v = ld4(x + (((long long)a_img * p.h + yy) * p.w + xx) * p.x_ld + c);
            av.x += wgt4[q] * v.x; av.y += wgt4[q] * v.y; av.z += wgt4[q] * v.z; av.w += wgt4[q] * v.w;
          }
        }
        av.x *= mask; av.y *= mask; av.z *= mask; av.w *= mask;
      }
    }
    float4 bv = b_valid ? ld4(b_ptr + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    As[a_k][a_row] = av.x; As[a_k + 1][a_row] = av.y; As[a_k + 2][a_row] = av.z; As[a_k + 3][a_row] = av.w;
    Bs[b_k][b_row] = bv.x; Bs[b_k + 1][b_row] = bv.y; Bs[b_k + 2][b_row] = bv.z; Bs[b_k + 3][b_row] = bv.w;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float ar[4] = {a.x, a.y, a.z, a.w}, br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int ox = (int)(m % wo), oy = (int)((m / wo) % ho);
    const int img = (int)(m / ((long long)wo * ho));
    const long long pix = (long long)oy * wo + ox;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co >= p.cout) continue;
      float v = acc[i][j];
      if (p.bias_map) v += __ldg(p.bias_map + pix * p.cout + co);
      if (p.coord_w) v += __ldg(p.coord_w + co) * (__fdiv_rn((float)ox, (float)(wo - 1)) * 2.f - 1.f) +
                          __ldg(p.coord_w + p.cout + co) * (__fdiv_rn((float)oy, (float)(ho - 1)) * 2.f - 1.f);
      v = v * __ldg(p.scale + co) + __ldg(p.shift + co);
      if (p.residual) {
        if (p.out_dtype == PPY_BF16) v += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.residual)[m * p.res_ld + co]);
        else v += reinterpret_cast<const float*>(p.residual)[m * p.res_ld + co];
      }
      v = apply_act(v, p.act);
      if (!p.upsample2x) {
        if (p.out_dtype == PPY_BF16) reinterpret_cast<__nv_bfloat16*>(p.y)[m * p.y_ld + co] = __float2bfloat16_rn(v);
        else reinterpret_cast<float*>(p.y)[m * p.y_ld + co] = v;
      } else {
        for (int q = 0; q < 4; ++q) {
          const long long dst = (((long long)img * 2 * ho + 2 * oy + (q >> 1)) * 2 * wo + 2 * ox + (q & 1)) * p.y_ld + co;
          if (p.out_dtype == PPY_BF16) reinterpret_cast<__nv_bfloat16*>(p.y)[dst] = __float2bfloat16_rn(v);
          else reinterpret_cast<float*>(p.y)[dst] = v;
        }
      }
    }
  }
}

}  // namespace

int validate_conv(const ppy_conv_params* p, int elem_bytes, int* ho, int* wo) {
  PPY_REQUIRE(p && p->x && p->weight && p->scale && p->shift && p->y);
  PPY_REQUIRE(p->n > 0 && p->h > 0 && p->w > 0 && p->cin > 0 && p->cout > 0);
  PPY_REQUIRE(p->kh > 0 && p->kw > 0 && p->kh == p->kw && p->stride >= 1 && p->pad >= 0);
  PPY_REQUIRE(p->cin % 8 == 0 && p->x_ld >= p->cin && p->y_ld >= p->cout);
  PPY_REQUIRE(p->k_pad % 64 == 0 && p->k_pad >= p->kh * p->kw * p->cin && p->cout_pad >= p->cout);
  PPY_REQUIRE((reinterpret_cast<uintptr_t>(p->x) & 15) == 0 && (p->x_ld * elem_bytes) % 16 == 0);
  PPY_REQUIRE((reinterpret_cast<uintptr_t>(p->weight) & 15) == 0);
  PPY_REQUIRE(p->act >= PPY_ACT_NONE && p->act <= PPY_ACT_MISH);
  PPY_REQUIRE(p->out_dtype == PPY_F32 || p->out_dtype == PPY_BF16 || p->out_dtype == PPY_F16X2);
  if (p->residual) PPY_REQUIRE(p->res_ld >= p->cout && !p->upsample2x);
  // the reference's DCN output size (H + 2p - (k-1)) // stride equals this for every shape it supports
  *ho = (p->h + 2 * p->pad - (p->kh - 1) - 1) / p->stride + 1;
  *wo = (p->w + 2 * p->pad - (p->kw - 1) - 1) / p->stride + 1;
  PPY_REQUIRE(*ho > 0 && *wo > 0);
  if (p->offset_mask) PPY_REQUIRE(p->om_ld >= 3 * p->kh * p->kw);
  if (p->coord_w) PPY_REQUIRE(p->kh == 1 && p->stride == 1 && p->pad == 0 && !p->bias_map && !p->accumulate && !p->upsample2x && *ho > 1 && *wo > 1);
  return PPY_OK;
}

}  // namespace ppy

extern "C" {
using namespace ppy;

int ppy_pack_conv_weight(const float* w_oihw, int cout, int cin_total, int kh, int kw, int c_begin, int c_count,
                         void* packed, int cout_pad, int cin_pad, int k_pad, int dtype, ppy_stream_t s) {
  PPY_REQUIRE(w_oihw && packed && cout > 0 && cin_total > 0 && kh > 0 && kw > 0);
  PPY_REQUIRE(c_begin >= 0 && c_count > 0 && c_begin + c_count <= cin_total && cin_pad >= c_count);
  PPY_REQUIRE(cout_pad >= cout && k_pad >= kh * kw * cin_pad);
  const long long total = (long long)cout_pad * k_pad;
  long long blocks = ceil_div(total, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  const dim3 grid2((unsigned)(k_pad >= 1024 ? 4 : 1), (unsigned)(cout_pad < 65535 ? cout_pad : 65535));      // (row, K chunk) blocks
  if (dtype == PPY_BF16)
    pack_weight_kernel<__nv_bfloat16><<<grid2, 256, 0, as_stream(s)>>>(
        w_oihw, cout, cin_total, kh, kw, c_begin, c_count, (__nv_bfloat16*)packed, cout_pad, cin_pad, k_pad);
  else if (dtype == PPY_F32)
    pack_weight_kernel<float><<<grid2, 256, 0, as_stream(s)>>>(w_oihw, cout, cin_total, kh, kw, c_begin,
                                                                c_count, (float*)packed, cout_pad, cin_pad, k_pad);
  else if (dtype == PPY_F16X2)
    pack_weight_pair_kernel<<<(unsigned)blocks, 256, 0, as_stream(s)>>>(w_oihw, cout, cin_total, kh, kw, c_begin, c_count,
                                                                         (__half*)packed, cout_pad, cin_pad, k_pad);
  else return PPY_ERR_INVALID;
  return check_launch();
}

int ppy_pack_conv_weight_dgrad(const float* w_oihw, int cout, int cin_total, int kh, int kw, int c_main, void* packed, int rows_pad, int o_pad,
                               int k_pad, int dtype, ppy_stream_t s) {
  PPY_REQUIRE(w_oihw && packed && cout > 0 && cin_total > 0 && kh > 0 && kw > 0 && c_main > 0 && c_main <= cin_total);
  PPY_REQUIRE(rows_pad >= c_main && o_pad >= cout && k_pad >= kh * kw * o_pad);
  if (dtype != PPY_BF16) return PPY_ERR_UNSUPPORTED;
  const int taps = kh * kw;
  dim3 grid((unsigned)ceil_div(o_pad, 32), (unsigned)ceil_div(rows_pad, 32), (unsigned)(taps + (k_pad > taps * o_pad ? 1 : 0)));
  pack_weight_dgrad_kernel<<<grid, 256, 0, as_stream(s)>>>(w_oihw, cout, cin_total, taps, c_main, (__nv_bfloat16*)packed, rows_pad, o_pad, k_pad);
  return check_launch();
}

int ppy_conv_f32(const ppy_conv_params* p, ppy_stream_t s) {
  int ho, wo;
  int rc = validate_conv(p, 4, &ho, &wo);
  if (rc) return rc;
  if (p->accumulate || p->split_k > 1 || p->wgrad_taps > 1 || p->x2_kb > 0) return PPY_ERR_UNSUPPORTED;   // tcgen05 path only
  PPY_REQUIRE(p->out_dtype == PPY_F32);
  const long long M = (long long)p->n * ho * wo;
  dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(p->cout, BN));
  const int k_true = p->kh * p->kw * p->cin;
  if (p->offset_mask) conv_simt_kernel<true><<<grid, THREADS, 0, as_stream(s)>>>(*p, ho, wo, k_true);
  else conv_simt_kernel<false><<<grid, THREADS, 0, as_stream(s)>>>(*p, ho, wo, k_true);
  return check_launch();
}

}  // extern "C"
