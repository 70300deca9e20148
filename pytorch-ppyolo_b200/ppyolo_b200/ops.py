"""Operator-level host API over the C ABI: torch CUDA tensors in, torch CUDA tensors out.

These are the functions the module surface in ``model/`` calls (one reference operator each).  They take
and return NCHW-shaped fp32 tensors like the reference's operators, convert to the kernels' NHWC layout
with the library's own layout kernels, and launch on torch's current stream.  There is no CPU path: a
non-CUDA tensor raises.  The whole-network fast path (``engine.InferenceEngine``) bypasses these
wrappers and drives the same C entry points over pre-allocated NHWC buffers.
"""
import ctypes
import weakref

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, ConvParams, PPY_F32, PPY_BF16, PPY_F16X2

_PRECISION = 'fp32'      # arithmetic of module-level convs: 'fp32' (SIMT) or 'bf16' (tcgen05)
PRECISIONS = ('bf16', 'fp32', 'f16x2')      # engine precisions; 'f16x2' = fp32-grade tensor-core path (fp16 hi/lo pairs)


def set_precision(p):
    global _PRECISION
    if p not in ('fp32', 'bf16'):
        raise ValueError(p)
    _PRECISION = p


def get_precision():
    return _PRECISION


def round_up(v, m):
    return (v + m - 1) // m * m


def _cuda(t, name='tensor'):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError('ppyolo_b200: %s must be a CUDA tensor -- the kernels have no CPU fallback' % name)
    return t


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def torch_dtype(code):
    return {PPY_BF16: torch.bfloat16, PPY_F16X2: torch.float16}.get(code, torch.float32)


def dtype_code(precision):
    return {'bf16': PPY_BF16, 'f16x2': PPY_F16X2}.get(precision, PPY_F32)


# ------------------------------------------------------------------------------------------------
# layout helpers (NCHW fp32 <-> NHWC)
# ------------------------------------------------------------------------------------------------
def to_nhwc(x, dtype_code_=PPY_F32, c_pad=None):
    """NCHW fp32 tensor -> new NHWC buffer [N,H,W,c_pad] (channels zero padded)."""
    _cuda(x, 'input')
    x = x.detach().float().contiguous()
    n, c, h, w = x.shape
    c_pad = c_pad or round_up(c, 8)
    y = torch.empty((n, h, w, c_pad), dtype=torch_dtype(dtype_code_), device=x.device)
    check(lib.ppy_nchw_to_nhwc(ptr(x), ptr(y), n, c, h, w, c_pad, dtype_code_, stream_ptr()), 'nchw_to_nhwc')
    return y


def from_nhwc(y, c):
    """NHWC buffer [N,H,W,ld] (fp32 or bf16) -> NCHW fp32 tensor with the first ``c`` channels."""
    n, h, w, ld = y.shape
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=y.device)
    code = PPY_BF16 if y.dtype == torch.bfloat16 else PPY_F32
    check(lib.ppy_nhwc_to_nchw(ptr(y), ld, code, ptr(out), n, c, h, w, stream_ptr()), 'nhwc_to_nchw')
    return out


# ------------------------------------------------------------------------------------------------
# conv + folded norm + act (+residual), DCNv2
# ------------------------------------------------------------------------------------------------
_PACK_CACHE = {}


def pack_weight(weight, code, c_begin=0, c_count=None, cache=True):
    """OIHW fp32 -> packed K-major [cout_pad, k_pad].

    Cached per live tensor object (weak reference + in-place version counter), so module parameters are
    packed once while temporaries can never alias a stale entry.  ``cache=False`` for weights updated behind torch's
    back (the fused SGD kernel writes parameters through raw pointers: the version counter does not move).
    """
    _cuda(weight, 'weight')
    cout, cin_total, kh, kw = weight.shape
    c_count = cin_total - c_begin if c_count is None else c_count
    key = (id(weight), code, c_begin, c_count)
    hit = _PACK_CACHE.get(key) if cache else None
    if hit is not None and hit[0]() is weight and hit[1] == weight._version and hit[2] == weight.data_ptr():
        return hit[3]
    cin_pad = round_up(c_count, 8)
    k_pad = round_up(kh * kw * cin_pad, 64)
    cout_pad = round_up(cout, 32)
    w32 = weight.detach().float().contiguous()
    packed = torch.empty((cout_pad, k_pad), dtype=torch_dtype(code), device=weight.device)
    check(lib.ppy_pack_conv_weight(ptr(w32), cout, cin_total, kh, kw, c_begin, c_count, ptr(packed), cout_pad, cin_pad,
                                   k_pad, code, stream_ptr()), 'pack_conv_weight')
    if len(_PACK_CACHE) > 4096:
        _PACK_CACHE.clear()
    result = (packed, cin_pad, k_pad, cout_pad)
    if not cache:
        return result
    _PACK_CACHE[key] = (weakref.ref(weight), weight._version, weight.data_ptr(), result)
    return result


def pack_weight_dgrad(weight, c_main, o_pad):
    """Packed bf16 weight of the input-gradient conv (dY with ``o_pad`` channels -> ``c_main`` channels) straight from the forward
    OIHW fp32 weight: 180-degree rotation + transposition + packing in one launch (ppy_pack_conv_weight_dgrad).  Returns the
    same tuple as ``pack_weight`` of the transposed weight."""
    _cuda(weight, 'weight')
    cout, cin_total, kh, kw = weight.shape
    w32 = weight.detach()
    if w32.dtype != torch.float32 or not w32.is_contiguous():
        w32 = w32.float().contiguous()
    cin_pad = round_up(o_pad, 8)
    k_pad = round_up(kh * kw * cin_pad, 64)
    rows_pad = round_up(c_main, 32)
    packed = torch.empty((rows_pad, k_pad), dtype=torch.bfloat16, device=weight.device)
    check(lib.ppy_pack_conv_weight_dgrad(ptr(w32), cout, cin_total, kh, kw, c_main, ptr(packed), rows_pad, cin_pad, k_pad, PPY_BF16,
                                         stream_ptr()), 'pack_conv_weight_dgrad')
    return packed, cin_pad, k_pad, rows_pad


def split_halves(t):
    """fp32 tensor -> (hi, lo) fp16 tensors with t = hi + lo to 2^-22 (torch ops; used for small host-built operands)."""
    hi = t.to(torch.float16)
    return hi, (t - hi.float()).to(torch.float16)


def coord_source(h, w, k, act_scale, device):
    """Batch-invariant second K source that carries CoordConv's two channels (model/custom_layers.py:256-272) into a k x k
    stride-1 conv of the pair path: pair tensor [2, rows, 64] with, per output pixel, the x / y coordinate each tap reads (zero
    where the tap falls into the padding) in columns 2*tap, 2*tap + 1.  rows = the image's h*w pixels repeated until any
    128-row tile that starts inside the image is contiguous.  Values are multiplied by ``act_scale`` like every pair activation."""
    xs = torch.arange(0, w, dtype=torch.float32, device=device) / (w - 1) * 2.0 - 1
    ys = torch.arange(0, h, dtype=torch.float32, device=device) / (h - 1) * 2.0 - 1
    src = torch.zeros((h, w, 64), dtype=torch.float32, device=device)
    r = (k - 1) // 2
    for ky in range(k):
        for kx in range(k):
            t = ky * k + kx
            y0, y1 = max(0, r - ky), min(h, h + r - ky)          # output rows whose tap row (oy + ky - r) is inside
            x0, x1 = max(0, r - kx), min(w, w + r - kx)
            src[y0:y1, x0:x1, 2 * t] = xs[x0 + kx - r:x1 + kx - r].view(1, -1)
            src[y0:y1, x0:x1, 2 * t + 1] = ys[y0 + ky - r:y1 + ky - r].view(-1, 1)
    src = src.reshape(h * w, 64) * act_scale
    reps = (h * w + 127 + h * w - 1) // (h * w)
    src = src.repeat(reps, 1)
    hi, lo = split_halves(src)
    return torch.stack([hi, lo]).contiguous()


def append_weight_block(packed, cout, extra):
    """[2, cout_pad, k_pad] packed pair weight + ``extra`` [cout, 64] fp32 (already multiplied by chan_scale) -> packed
    [2, cout_pad, k_pad + 64]."""
    _, cout_pad, k_pad = packed.shape
    out = torch.zeros((2, cout_pad, k_pad + 64), dtype=torch.float16, device=packed.device)
    out[:, :, :k_pad] = packed
    hi, lo = split_halves(extra.float())
    out[0, :cout, k_pad:], out[1, :cout, k_pad:] = hi, lo
    return out.contiguous()


def pack_weight_pair(weight, c_begin=0, c_count=None, amax_with=None):
    """OIHW fp32 -> PPY_F16X2 packing [2, cout_pad, k_pad] fp16 (hi plane, lo plane) for ppy_conv_f16x2.

    Every output channel is first multiplied by a power of two ``chan_scale[co]`` that brings its largest weight into
    [2^13, 2^14), so hi AND lo parts of all but vanishing weights are normal fp16 numbers (22 significant bits); the
    caller divides the channel's epilogue scale by ``chan_scale`` (exact).  Returns (packed, cin_pad, k_pad, cout_pad,
    chan_scale[cout] fp32)."""
    _cuda(weight, 'weight')
    cout, cin_total, kh, kw = weight.shape
    c_count = cin_total - c_begin if c_count is None else c_count
    cin_pad = round_up(c_count, 8)
    k_pad = round_up(kh * kw * cin_pad, 64)
    cout_pad = round_up(cout, 32)
    w32 = weight.detach().float()
    amax = w32[:, c_begin:c_begin + c_count].abs().amax(dim=(1, 2, 3))
    if amax_with is not None:                       # further weights of the same rows that will share chan_scale
        amax = torch.maximum(amax, amax_with.float())
    _, ex = torch.frexp(amax)                       # amax = m * 2^ex, m in [0.5, 1)
    chan_scale = torch.where(amax > 0, torch.ldexp(torch.ones_like(amax), (14 - ex).clamp(-30, 40)), torch.ones_like(amax))
    w32 = (w32 * chan_scale.view(-1, 1, 1, 1)).contiguous()
    packed = torch.empty((2, cout_pad, k_pad), dtype=torch.float16, device=weight.device)
    check(lib.ppy_pack_conv_weight(ptr(w32), cout, cin_total, kh, kw, c_begin, c_count, ptr(packed), cout_pad, cin_pad,
                                   k_pad, PPY_F16X2, stream_ptr()), 'pack_conv_weight')
    return packed, cin_pad, k_pad, cout_pad, chan_scale.contiguous()


def ident_tile(cout, cin):
    """N tile of ppy_conv_f16x2 for a 1x1 conv whose identity shortcut runs as K blocks (ppy_conv_params.x2_tiled): 256 when the
    whole K (cin + 256) fits one accumulator (<= 8 blocks), else 128; None = shape not supported."""
    if cin % 64:
        return None
    if cout % 256 == 0 and cin // 64 + 4 <= 8:
        return 256
    return 128 if cout % 128 == 0 else None


def pack_weight_pair_dual(w1, scale1, w2=None, scale2=None, ident_bn=None):
    """Packed PPY_F16X2 weight [2, cout_pad, cin + extra] of a 1x1 conv with a second K source (ppy_conv_params.x2).

    ``w2`` given: K-concatenation [w1 * scale1 | w2 * scale2] (two 1x1 convs with their folded norm scales, one GEMM).
    ``ident_bn`` given: identity shortcut -- ``ident_bn`` extra columns per row holding ``chan_scale[co]`` at column
    ``co % ident_bn`` (a power of two, exact in fp16, lo part zero).  Rows are multiplied by the power of two ``chan_scale``
    that brings the largest entry into [2^13, 2^14) (identity mode: capped at 2^15 so the diagonal stays representable).
    Returns (packed, k_pad, cout_pad, chan_scale) or None when a channel's weights are too small for the cap."""
    _cuda(w1, 'weight')
    cout, cin = w1.shape[0], w1.shape[1]
    assert w1.shape[2] == 1 and w1.shape[3] == 1 and cin % 64 == 0
    wa = w1.detach().float().reshape(cout, cin) * scale1.detach().float().view(-1, 1)
    if w2 is not None:
        wb = w2.detach().float().reshape(cout, -1) * scale2.detach().float().view(-1, 1)
        assert wb.shape[1] % 64 == 0
        amax = torch.cat([wa, wb], 1).abs().amax(dim=1)
        _, ex = torch.frexp(amax)
        chan_scale = torch.where(amax > 0, torch.ldexp(torch.ones_like(amax), (14 - ex).clamp(-30, 40)), torch.ones_like(amax))
    else:
        amax = wa.abs().amax(dim=1)
        _, ex = torch.frexp(amax)
        chan_scale = torch.where(amax > 0, torch.ldexp(torch.ones_like(amax), (14 - ex).clamp(-14, 15)), torch.ones_like(amax))
        if bool(((amax * chan_scale < 1.0) & (amax > 0)).any()):
            return None
        wb = torch.zeros((cout, ident_bn), dtype=torch.float32, device=w1.device)
        idx = torch.arange(cout, device=w1.device)
        wb[idx, idx % ident_bn] = 1.0
    w = (torch.cat([wa, wb], 1) * chan_scale.view(-1, 1)).contiguous()
    k_pad = w.shape[1]
    cout_pad = round_up(cout, 32)
    packed = torch.empty((2, cout_pad, k_pad), dtype=torch.float16, device=w1.device)
    check(lib.ppy_pack_conv_weight(ptr(w), cout, k_pad, 1, 1, 0, k_pad, ptr(packed), cout_pad, k_pad, k_pad, PPY_F16X2, stream_ptr()),
          'pack_conv_weight')
    return packed, k_pad, cout_pad, chan_scale.contiguous()


def conv_pair_dual(x, w1, scale1, shift1, x2, w2=None, scale2=None, shift2=None, act=0, overflow=None):
    """1x1 conv of the f16x2 path with a second K source (stand-alone wrapper; the engine builds the same parameters):
    ``w2`` given -> act(scale1*conv(x, w1) + shift1 + scale2*conv(x2, w2) + shift2) as one GEMM; ``w2`` None -> x2 is the
    identity shortcut, act(scale1*conv(x, w1) + shift1 + x2) with the add done by the tensor core."""
    _cuda(x, 'input')
    cout, cin = w1.shape[0], w1.shape[1]
    bn = None
    if w2 is None:
        bn = ident_tile(cout, cin)
        if bn is None:
            raise ValueError('identity K blocks need cin %% 64 == 0 and cout %% 128 == 0')
    res = pack_weight_pair_dual(w1, scale1, w2, scale2, bn)
    if res is None:
        raise ValueError('weights too small for the identity-block scaling')
    packed, k_pad, cout_pad, cs = res
    _, n, h, w, ld = x.shape
    ldo = round_up(cout, 8)
    out = torch.zeros((2, n, h, w, ldo), dtype=torch.float16, device=x.device)
    sc = (1.0 / cs).contiguous()
    sh = _f32(shift1).to(x.device)
    if shift2 is not None:
        sh = (sh + _f32(shift2).to(x.device)).contiguous()
    p = ConvParams()
    p.x, p.x_ld, p.x_plane = x.data_ptr(), ld, x.stride(0)
    p.n, p.h, p.w, p.cin = n, h, w, cin
    p.weight = packed.data_ptr()
    p.cout, p.kh, p.kw, p.stride, p.pad = cout, 1, 1, 1, 0
    p.k_pad, p.cout_pad = k_pad, cout_pad
    p.scale, p.shift = sc.data_ptr(), sh.data_ptr()
    p.act = act
    p.y, p.y_ld, p.out_dtype, p.y_plane = out.data_ptr(), ldo, PPY_F16X2, out.stride(0)
    p.x2, p.x2_ld, p.x2_plane = x2.data_ptr(), x2.shape[-1], x2.stride(0)
    p.x2_kb, p.x2_tiled = (k_pad - cin) // 64, 1 if w2 is None else 0
    p.overflow = overflow.data_ptr() if overflow is not None else None
    check(lib.ppy_conv_f16x2(ctypes.byref(p), stream_ptr()), 'conv_f16x2')
    return out


def split_pair(x_nhwc):
    """NHWC fp32 tensor -> PPY_F16X2 tensor [2, N, H, W, C] (hi plane, lo plane)."""
    _cuda(x_nhwc)
    x = x_nhwc.detach().float().contiguous()
    n, h, w, c = x.shape
    y = torch.empty((2, n, h, w, c), dtype=torch.float16, device=x.device)
    check(lib.ppy_split_f16x2(ptr(x), c, ptr(y), c, y.stride(0), n * h * w, c, stream_ptr()), 'split_f16x2')
    return y


def join_pair(y, c=None):
    """PPY_F16X2 tensor [2, N, H, W, ld] -> NHWC fp32 tensor with the first ``c`` channels."""
    _, n, h, w, ld = y.shape
    c = ld if c is None else c
    out = torch.empty((n, h, w, c), dtype=torch.float32, device=y.device)
    check(lib.ppy_join_f16x2(ptr(y), ld, y.stride(0), ptr(out), c, n * h * w, c, stream_ptr()), 'join_f16x2')
    return out


def conv_nhwc(x, packed, cin, cout, k, stride, pad, scale, shift, act, code, residual=None, bias_map=None,
              out=None, out_code=None, upsample2x=False, offset_mask=None, x_ld=None, coord_w=None, accumulate=False, split_k=0):
    """Launch one fused conv on NHWC buffers. ``x`` [N,H,W,ld]; returns ``out`` [N,Ho,Wo,ld_out]."""
    w_packed, cin_pad, k_pad, cout_pad = packed
    n, h, w, ld = x.shape
    ho = (h + 2 * pad - k) // stride + 1
    wo = (w + 2 * pad - k) // stride + 1
    out_code = code if out_code is None else out_code
    if out is None:
        oh, ow = (2 * ho, 2 * wo) if upsample2x else (ho, wo)
        out = torch.empty((n, oh, ow, round_up(cout, 8)), dtype=torch_dtype(out_code), device=x.device)
    p = ConvParams()
    p.x, p.x_ld = x.data_ptr(), (x_ld or ld)
    p.n, p.h, p.w, p.cin = n, h, w, cin_pad
    p.weight = w_packed.data_ptr()
    p.cout, p.kh, p.kw, p.stride, p.pad = cout, k, k, stride, pad
    p.k_pad, p.cout_pad = k_pad, cout_pad
    p.scale, p.shift = scale.data_ptr(), shift.data_ptr()
    p.bias_map = bias_map.data_ptr() if bias_map is not None else None
    p.coord_w = coord_w.data_ptr() if coord_w is not None else None
    p.accumulate, p.split_k = (1 if accumulate else 0), split_k      # partial sums are ADDED into a caller-zeroed fp32 `out`
    p.residual = residual.data_ptr() if residual is not None else None
    p.res_ld = residual.shape[-1] if residual is not None else 0
    # Mish (reference model/custom_layers.py:37-43; no PP-YOLO config uses it): the tcgen05 epilogue's activation is a slope
    # select, so on that path the conv leaves the pre-activation and ppy_activation finishes it in place
    mish_pass = act == _lib.ACT_MISH and code == PPY_BF16
    p.act = 0 if mish_pass else act
    p.y, p.y_ld, p.out_dtype = out.data_ptr(), out.shape[-1], out_code
    p.upsample2x = 1 if upsample2x else 0
    p.offset_mask = offset_mask.data_ptr() if offset_mask is not None else None
    p.om_ld = offset_mask.shape[-1] if offset_mask is not None else 0
    fn = lib.ppy_conv_bf16 if code == PPY_BF16 else lib.ppy_conv_f32
    check(fn(ctypes.byref(p), stream_ptr()), 'conv_bf16' if code == PPY_BF16 else 'conv_f32')
    if mish_pass:
        check(lib.ppy_activation(out.data_ptr(), out.numel(), _lib.ACT_MISH, out_code, stream_ptr()), 'activation')
    return out


def conv_pair(x, weight, scale, shift, stride=1, pad=0, act=0, residual=None, bias_map=None, out_f32=False,
              upsample2x=False, offset_mask=None, c_count=None, overflow=None, coord=False):
    """One fused conv of the fp32-grade tensor-core path (ppy_conv_f16x2): ``x`` / ``residual`` are PPY_F16X2 tensors
    [2, N, H, W, ld] (fp16 hi | lo planes), ``weight`` the OIHW fp32 tensor (its first ``c_count`` input channels are
    used).  Returns a pair tensor [2, N, Ho, Wo, ld_out], or an NHWC fp32 tensor with ``out_f32``."""
    _cuda(x, 'input')
    assert x.dtype == torch.float16 and x.dim() == 5 and x.shape[0] == 2
    cout, cin_total, k, _ = weight.shape
    x2 = None
    if coord:        # the weight's channels [c_count, c_count + 2) are CoordConv's: one extra K block from a batch-invariant source
        wc = weight.detach().float()[:, c_count:c_count + 2]
        packed, cin_pad, k_pad, cout_pad, cs = pack_weight_pair(weight, 0, c_count, amax_with=wc.abs().amax(dim=(1, 2, 3)))
        extra = torch.zeros((cout, 64), dtype=torch.float32, device=x.device)
        extra[:, :2 * k * k] = wc.permute(0, 2, 3, 1).reshape(cout, 2 * k * k)
        packed = append_weight_block(packed, cout, extra * cs.view(-1, 1))
        k_pad += 64
        x2 = coord_source(x.shape[2], x.shape[3], k, 1.0, x.device)
    else:
        packed, cin_pad, k_pad, cout_pad, cs = pack_weight_pair(weight, 0, c_count)
    _, n, h, w, ld = x.shape
    ho = (h + 2 * pad - k) // stride + 1
    wo = (w + 2 * pad - k) // stride + 1
    oh, ow = (2 * ho, 2 * wo) if upsample2x else (ho, wo)
    ldo = round_up(cout, 8)
    if out_f32:
        out = torch.zeros((n, oh, ow, ldo), dtype=torch.float32, device=x.device)
    else:
        out = torch.zeros((2, n, oh, ow, ldo), dtype=torch.float16, device=x.device)
    sc = (_f32(scale).to(x.device) / cs).contiguous()
    sh = _f32(shift).to(x.device)
    bm = (bias_map.float().to(x.device) * cs).contiguous() if bias_map is not None else None
    p = ConvParams()
    p.x, p.x_ld, p.x_plane = x.data_ptr(), ld, x.stride(0)
    p.n, p.h, p.w, p.cin = n, h, w, cin_pad
    p.weight = packed.data_ptr()
    p.cout, p.kh, p.kw, p.stride, p.pad = cout, k, k, stride, pad
    p.k_pad, p.cout_pad = k_pad, cout_pad
    p.scale, p.shift = sc.data_ptr(), sh.data_ptr()
    p.bias_map = bm.data_ptr() if bm is not None else None
    p.residual = residual.data_ptr() if residual is not None else None
    p.res_ld = residual.shape[-1] if residual is not None else 0
    p.res_plane = residual.stride(0) if residual is not None else 0
    p.act = act
    p.y, p.y_ld = out.data_ptr(), ldo
    p.out_dtype = PPY_F32 if out_f32 else PPY_F16X2
    p.y_plane = 0 if out_f32 else out.stride(0)
    p.upsample2x = 1 if upsample2x else 0
    p.offset_mask = offset_mask.data_ptr() if offset_mask is not None else None
    p.om_ld = offset_mask.shape[-1] if offset_mask is not None else 0
    p.overflow = overflow.data_ptr() if overflow is not None else None
    if x2 is not None:
        p.x2, p.x2_ld, p.x2_plane = x2.data_ptr(), 64, x2.stride(0)
        p.x2_kb, p.x2_tiled, p.x2_row_mod, p.x2_rows = 1, 0, h * w, x2.shape[1]
    check(lib.ppy_conv_f16x2(ctypes.byref(p), stream_ptr()), 'conv_f16x2')
    return out


def _f32(t):
    return t.detach().float().contiguous()


def conv_bn_act(x, weight, scale, shift, stride=1, padding=0, act=0, residual=None, precision=None,
                keep_nhwc=False):
    """Conv2dUnit.forward (reference model/custom_layers.py:243-253) as one kernel; NCHW fp32 in/out."""
    precision = precision or _PRECISION
    code = dtype_code(precision)
    cout, cin, k, _ = weight.shape
    packed = pack_weight(weight, code)
    xh = to_nhwc(x, code, packed[1])
    res = to_nhwc(residual, code) if residual is not None else None
    y = conv_nhwc(xh, packed, cin, cout, k, stride, padding, _f32(scale), _f32(shift), act, code, residual=res)
    return y if keep_nhwc else from_nhwc(y, cout)


def conv_unit_residual_relu(unit, x, shortcut):
    """Last conv of a residual block with the block's `x + shortcut; relu` fused into its epilogue
    (reference model/resnet_vd.py:54-56, :84-86, :264-266). The unit itself has act=None."""
    scale, shift = unit.folded_scale_shift()
    if unit.bn is not None and unit.bn.training:
        raise NotImplementedError('train-mode BatchNorm is not built yet; call model.eval() first')
    return conv_bn_act(x, unit.conv.weight, scale, shift, stride=unit.stride, padding=unit.padding,
                       act=_lib.ACT_RELU, residual=shortcut)


def dcnv2(x, offset_w, offset_b, dcn_w, dcn_b=None, stride=1, padding=1, scale=None, shift=None, act=0,
          residual=None, precision=None):
    """DCNv2.forward (reference model/custom_layers.py:551-677) + optional folded norm/act epilogue."""
    precision = precision or _PRECISION
    code = dtype_code(precision)
    cout, cin, k, _ = dcn_w.shape
    dev = x.device
    xh = to_nhwc(x, code)
    # offset/mask conv: plain conv with bias, fp32 output kept NHWC for the sampler
    om_packed = pack_weight(offset_w, code)
    n_om = offset_w.shape[0]
    ones = torch.ones(n_om, dtype=torch.float32, device=dev)
    om = conv_nhwc(xh, om_packed, cin, n_om, k, stride, padding, ones, _f32(offset_b), 0, code, out_code=PPY_F32)
    packed = pack_weight(dcn_w, code)
    if scale is None:
        scale = torch.ones(cout, dtype=torch.float32, device=dev)
    if shift is None:
        shift = torch.zeros(cout, dtype=torch.float32, device=dev)
    if dcn_b is not None:
        shift = shift + _f32(dcn_b) * scale
    res = to_nhwc(residual, code) if residual is not None else None
    y = conv_nhwc(xh, packed, cin, cout, k, stride, padding, _f32(scale), _f32(shift), act, code, residual=res,
                  offset_mask=om)
    return from_nhwc(y, cout)


# ------------------------------------------------------------------------------------------------
# glue ops
# ------------------------------------------------------------------------------------------------
def _pool(fn, x, out_hw, name):
    xh = to_nhwc(x)
    n, h, w, ld = xh.shape
    y = torch.empty((n, out_hw[0], out_hw[1], ld), dtype=torch.float32, device=x.device)
    check(fn(ptr(xh), ld, ptr(y), ld, n, h, w, ld, PPY_F32, stream_ptr()), name)
    return from_nhwc(y, x.shape[1])


def max_pool3s2(x):
    """MaxPool2d(3, 2, 1), reference model/resnet_vd.py:103."""
    return _pool(lib.ppy_maxpool3x3s2, x, ((x.shape[2] + 1) // 2, (x.shape[3] + 1) // 2), 'maxpool3x3s2')


def avg_pool2(x):
    """AvgPool2d(2, 2, 0), reference model/resnet_vd.py:30."""
    return _pool(lib.ppy_avgpool2x2, x, (x.shape[2] // 2, x.shape[3] // 2), 'avgpool2x2')


def spp(x, descending=False):
    """SPP (reference model/custom_layers.py:275-290)."""
    xh = to_nhwc(x)
    n, h, w, ld = xh.shape
    c = x.shape[1]
    if ld != c:
        raise NotImplementedError('SPP needs a channel count that is a multiple of 8')
    y = torch.empty((n, h, w, 4 * c), dtype=torch.float32, device=x.device)
    check(lib.ppy_spp(ptr(xh), ld, ptr(y), 4 * c, n, h, w, c, PPY_F32, stream_ptr()), 'spp')
    out = from_nhwc(y, 4 * c)
    if descending:
        out = torch.cat([out[:, 3 * c:], out[:, 2 * c:3 * c], out[:, c:2 * c], out[:, :c]], dim=1)
    return out


def coord_concat(x):
    """CoordConv (reference model/custom_layers.py:256-272): cat([x, xs, ys], dim=1)."""
    _cuda(x, 'input')
    n, c, h, w = x.shape
    coords = torch.empty((1, h, w, 8), dtype=torch.float32, device=x.device)
    check(lib.ppy_coord_channels(ptr(coords), 8, h, w, PPY_F32, stream_ptr()), 'coord_channels')
    cc = from_nhwc(coords, 2).expand(n, 2, h, w)
    return torch.cat([x, cc], dim=1)


def upsample2x_concat(route, feat):
    """nearest x2 upsample of ``route`` then channel concat with ``feat`` (reference model/head.py:390-397)."""
    rh, fh = to_nhwc(route), to_nhwc(feat)
    n, h, w, rc = rh.shape
    fc = fh.shape[-1]
    if rc != route.shape[1] or fc != feat.shape[1]:
        raise NotImplementedError('concat needs channel counts that are multiples of 8')
    y = torch.empty((n, 2 * h, 2 * w, rc + fc), dtype=torch.float32, device=route.device)
    check(lib.ppy_upsample2x(ptr(rh), rc, ptr(y), rc + fc, n, h, w, rc, PPY_F32, stream_ptr()), 'upsample2x')
    ysl = ctypes.c_void_p(y.data_ptr() + 4 * rc)
    check(lib.ppy_copy_channels(ptr(fh), fc, ysl, rc + fc, n * 4 * h * w, fc, PPY_F32, stream_ptr()), 'copy_channels')
    return from_nhwc(y, rc + fc)


def activation(x, act):
    y = _cuda(x).detach().float().contiguous().clone()
    check(lib.ppy_activation(ptr(y), y.numel(), act, PPY_F32, stream_ptr()), 'activation')
    return y


# ------------------------------------------------------------------------------------------------
# DropBlock (training)
# ------------------------------------------------------------------------------------------------
_DROPBLOCK_RNG = {}


def dropblock_rng(device):
    """Device-resident (seed, offset) pair of the DropBlock Philox streams (int64[2]); advanced by the kernels."""
    st = _DROPBLOCK_RNG.get(device)
    if st is None:
        st = torch.tensor([torch.initial_seed() & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64, device=device)
        _DROPBLOCK_RNG[device] = st
    return st


def dropblock_seed(seed, device='cuda'):
    device = torch.device(device) if not isinstance(device, torch.device) else device
    if device.index is None:
        device = torch.device('cuda', torch.cuda.current_device())
    dropblock_rng(device).copy_(torch.tensor([int(seed), 0], dtype=torch.int64))


def dropblock_gamma(h, block_size, keep_prob):
    """CalculateGamma, model/custom_layers.py:307-323 (fp32 tensor arithmetic in the reference)."""
    f = np.float32
    return float(f(f(h) * f(h)) * f(1 - keep_prob) / (f(block_size * block_size) * f((h - block_size + 1) ** 2)))


def _dense_like(x):
    return x.is_contiguous() or x.is_contiguous(memory_format=torch.channels_last)


class _DropBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, block_size, keep_prob, seeds):
        _cuda(x, 'input')
        if x.dtype not in (torch.float32, torch.bfloat16) or x.dim() != 4:
            raise TypeError('drop_block: fp32 / bf16 NCHW-shaped tensors only')
        if not _dense_like(x):
            x = x.contiguous()
        n, c, h, w = x.shape
        mask = torch.empty_like(x, dtype=torch.uint8)            # preserves x's memory format
        count = torch.empty(1, dtype=torch.int32, device=x.device)
        rng = dropblock_rng(x.device)
        st = [int(v) for v in x.stride()]
        if seeds is None:
            sd = torch.empty_like(mask)
            check(lib.ppy_dropblock_mask(ptr(sd), ptr(mask), n, c, h, w, st[0], st[1], st[2], st[3], int(block_size),
                                         dropblock_gamma(h, block_size, keep_prob), ptr(rng), ptr(count), stream_ptr()), 'dropblock_mask')
        else:
            sd = torch.empty_like(mask)
            sd.copy_(seeds.to(torch.uint8))
            check(lib.ppy_dropblock_mask_from_seeds(ptr(sd), ptr(mask), n, c, h, w, st[0], st[1], st[2], st[3], ptr(rng), ptr(count),
                                                    stream_ptr()), 'dropblock_mask_from_seeds')
        y = torch.empty_like(x)
        code = PPY_BF16 if x.dtype == torch.bfloat16 else PPY_F32
        check(lib.ppy_dropblock_apply(ptr(x), ptr(y), ptr(mask), ptr(count), x.numel(), code, stream_ptr()), 'dropblock_apply')
        ctx.save_for_backward(mask, count)
        return y

    @staticmethod
    def backward(ctx, dy):
        mask, count = ctx.saved_tensors
        if dy.stride() != mask.stride():
            dy = dy.contiguous(memory_format=torch.channels_last) if mask.is_contiguous(memory_format=torch.channels_last) and not mask.is_contiguous() else dy.contiguous()
        dx = torch.empty_like(dy)
        code = PPY_BF16 if dy.dtype == torch.bfloat16 else PPY_F32
        check(lib.ppy_dropblock_apply(ptr(dy), ptr(dx), ptr(mask), ptr(count), dy.numel(), code, stream_ptr()), 'dropblock_apply')
        return dx, None, None, None


def drop_block(x, block_size=3, keep_prob=0.9, seeds=None):
    """DropBlock.__call__ in training mode (reference model/custom_layers.py:303-342); differentiable.  ``seeds``: optional
    0/1 tensor of x's shape replacing the Bernoulli draw (parity tests inject the reference's own draw)."""
    return _DropBlockFn.apply(x, block_size, keep_prob, seeds)


# ------------------------------------------------------------------------------------------------
# pre-processing
# ------------------------------------------------------------------------------------------------
def pack_images(images, pinned=True):
    """List of HWC uint8 3-channel images (any sizes) -> (blob uint8 [total bytes], meta int64 [N, 3] = offset, h, w) host tensors
    (pinned for an asynchronous upload)."""
    metas, total = [], 0
    for im in images:
        if im.dtype != np.uint8 or im.ndim != 3 or im.shape[2] != 3:
            raise ValueError('pack_images: HWC uint8 images with 3 channels expected, got %s %s' % (im.dtype, im.shape))
        metas.append((total, im.shape[0], im.shape[1]))
        total += (im.size + 15) // 16 * 16
    blob = torch.empty(total, dtype=torch.uint8, pin_memory=pinned and torch.cuda.is_available())
    view = blob.numpy()
    for (off, h, w), im in zip(metas, images):
        view[off:off + im.size] = np.ascontiguousarray(im).reshape(-1)
    meta = torch.tensor(metas, dtype=torch.int64)
    if pinned and torch.cuda.is_available():
        meta = meta.pin_memory()
    return blob, meta


def resize_cubic_u8(blob, meta, size, swap_rb=True, out=None):
    """cv2.resize(img, fx=size/w, fy=size/h, interpolation=cv2.INTER_CUBIC) (+ BGR->RGB) of every packed image on the GPU:
    device ``blob`` / ``meta`` from ``pack_images`` -> uint8 [N, size, size, 3]."""
    _cuda(blob, 'image blob')
    _cuda(meta, 'image meta')
    n = meta.shape[0]
    if out is None:
        out = torch.empty((n, size, size, 3), dtype=torch.uint8, device=blob.device)
    check(lib.ppy_resize_cubic_u8_batch(ptr(blob), ptr(meta), n, ptr(out), int(size), 1 if swap_rb else 0, stream_ptr()), 'resize_cubic_u8')
    return out


# ------------------------------------------------------------------------------------------------
# training: fused loss + target assignment
# ------------------------------------------------------------------------------------------------
LOSS_NAMES = ('loss_xy', 'loss_wh', 'loss_obj', 'loss_cls', 'loss_iou', 'loss_iou_aware')


class _YoloLossFn(torch.autograd.Function):
    """All scales of the fine-grained YOLOv3 loss (reference model/losses.py:121-356) on the fused kernels: forward = one launch
    per scale adding into a 6-vector of losses, backward = one launch per scale writing d(sum_k g_k loss_k)/d(output)."""

    @staticmethod
    def forward(ctx, meta, gt_box, *tensors):
        ns = len(meta['scales'])
        outs = [t.detach().float().contiguous() for t in tensors[:ns]]
        tgts = [t.detach().float().contiguous() for t in tensors[ns:]]
        gt = gt_box.detach().float().contiguous()
        dev = outs[0].device
        losses = torch.zeros(6, dtype=torch.float32, device=dev)
        masks = []
        for o, t, sc in zip(outs, tgts, meta['scales']):
            n, _, size, _ = o.shape
            anchors = np.ascontiguousarray(np.asarray(sc['anchors'], dtype=np.float32).reshape(-1))
            a = anchors.size // 2
            per = 5 + meta['num_classes'] + (1 if meta['iou_aware'] else 0)
            if o.shape[1] != a * per or tuple(t.shape) != (n, a, 6 + meta['num_classes'], size, size):
                raise ValueError('yolo_loss: output %s / target %s do not match %d anchors x %d channels' % (tuple(o.shape), tuple(t.shape), a, per))
            noobj = torch.empty((n, a, size, size), dtype=torch.float32, device=dev)
            ws = torch.zeros(int(lib.ppy_yolo_loss_workspace_bytes(n, a, size)), dtype=torch.uint8, device=dev)
            check(lib.ppy_yolo_loss_forward(ptr(o), ptr(t), ptr(gt), n, a, meta['num_classes'], size, gt.shape[1],
                                            anchors.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), int(sc['stride']), float(sc['scale_x_y']),
                                            float(meta['ignore_thresh']), 1 if meta['iou_aware'] else 0, 1 if meta['has_iou_loss'] else 0,
                                            float(meta['iou_w']), 1 if meta['loss_square'] else 0, float(meta['aware_w']),
                                            1 if meta['match_score'] else 0, ptr(noobj), ptr(ws), ptr(losses), stream_ptr()), 'yolo_loss_forward')
            masks.append(noobj)
        ctx.meta = meta
        ctx.save_for_backward(*(outs + tgts + masks))
        return losses

    @staticmethod
    def backward(ctx, g):
        meta = ctx.meta
        ns = len(meta['scales'])
        saved = ctx.saved_tensors
        outs, tgts, masks = saved[:ns], saved[ns:2 * ns], saved[2 * ns:]
        g = g.detach().float().contiguous()
        grads = []
        for o, t, m, sc in zip(outs, tgts, masks, meta['scales']):
            n, _, size, _ = o.shape
            anchors = np.ascontiguousarray(np.asarray(sc['anchors'], dtype=np.float32).reshape(-1))
            a = anchors.size // 2
            go = torch.empty_like(o)
            check(lib.ppy_yolo_loss_backward(ptr(o), ptr(t), n, a, meta['num_classes'], size,
                                             anchors.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), int(sc['stride']), float(sc['scale_x_y']),
                                             1 if meta['iou_aware'] else 0, 1 if meta['has_iou_loss'] else 0, float(meta['iou_w']),
                                             1 if meta['loss_square'] else 0, float(meta['aware_w']), ptr(m), ptr(g), ptr(go), stream_ptr()),
                  'yolo_loss_backward')
            grads.append(go)
        return (None, None) + tuple(grads) + (None,) * ns


def yolo_loss_fused(outputs, targets, gt_box, scales, num_classes, ignore_thresh, iou_aware, has_iou_loss, iou_w, loss_square, aware_w,
                    match_score):
    """Six-vector (LOSS_NAMES order) of the fine-grained YOLOv3 loss summed over ``outputs`` (raw head outputs, NCHW) -- fused
    forward/backward kernels, differentiable w.r.t. the outputs.  ``scales``: per output dict(anchors=[w0,h0,...], stride, scale_x_y)."""
    for t in list(outputs) + list(targets) + [gt_box]:
        _cuda(t, 'yolo_loss input')
    meta = dict(scales=list(scales), num_classes=int(num_classes), ignore_thresh=ignore_thresh, iou_aware=bool(iou_aware),
                has_iou_loss=bool(has_iou_loss), iou_w=iou_w, loss_square=bool(loss_square), aware_w=aware_w, match_score=bool(match_score))
    return _YoloLossFn.apply(meta, gt_box, *(list(outputs) + list(targets)))


def gt2yolo_target_gpu(gt_bbox, gt_class, gt_score, anchors, anchor_masks, downsample_ratios, num_classes, h, w, iou_thresh=1.):
    """Gt2YoloTargetSingle (reference tools/transform.py:1318-1421) for a whole padded batch on the GPU: gt_bbox [N,G,4] normalised
    (cx,cy,w,h), gt_class [N,G] int, gt_score [N,G] (device tensors) -> list of [N, A, 6+C, h/s, w/s] fp32 device tensors."""
    _cuda(gt_bbox, 'gt_bbox')
    gb = gt_bbox.detach().float().contiguous()
    gc = gt_class.detach().to(torch.int32).contiguous()
    gs = gt_score.detach().float().contiguous()
    n, g = gb.shape[0], gb.shape[1]
    flat = np.ascontiguousarray(np.asarray(anchors, dtype=np.int32).reshape(-1))
    if not np.array_equal(flat, np.asarray(anchors).reshape(-1)):
        raise ValueError('gt2yolo_target_gpu: anchors must be whole pixels (the reference configs use ints)')
    out = []
    for mask, ratio in zip(anchor_masks, downsample_ratios):
        m = np.ascontiguousarray(np.asarray(mask, dtype=np.int32))
        gh_, gw_ = int(h / ratio), int(w / ratio)
        if gh_ * ratio != h or gw_ * ratio != w:
            raise ValueError('gt2yolo_target_gpu: image size must be a multiple of the downsample ratio')
        t = torch.empty((n, len(mask), 6 + num_classes, gh_, gw_), dtype=torch.float32, device=gb.device)
        check(lib.ppy_gt2yolo_target(ptr(gb), ptr(gc), ptr(gs), n, g, flat.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), flat.size // 2,
                                     m.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), m.size, int(num_classes), int(h), int(w), int(ratio),
                                     float(iou_thresh), ptr(t), stream_ptr()), 'gt2yolo_target')
        out.append(t)
    return out


# ------------------------------------------------------------------------------------------------
# head post-processing
# ------------------------------------------------------------------------------------------------
def iou_aware_score(output, an_num, num_classes, factor):
    xh = to_nhwc(output, PPY_F32, output.shape[1])
    n, h, w, ld = xh.shape
    oc = an_num * (5 + num_classes)
    y = torch.empty((n, h, w, oc), dtype=torch.float32, device=output.device)
    check(lib.ppy_iou_aware_score(ptr(xh), ld, ptr(y), oc, n * h * w, an_num, num_classes, float(factor), stream_ptr()),
          'iou_aware_score')
    return from_nhwc(y, oc)


def yolo_decode_nhwc(head, ld, n, size, anchors, stride, num_classes, scale_x_y, im_size, clip_bbox, iou_aware,
                     factor, boxes, scores, box_offset, total_boxes, hist_threshold=None, nms_workspace=None):
    """``nms_workspace`` + ``hist_threshold``: also count every score > threshold in the workspace's score histogram
    (consumed by ``matrix_nms_launch(..., have_hist=True)``)."""
    anchors = np.ascontiguousarray(np.asarray(anchors, dtype=np.float32).reshape(-1))
    an = anchors.size // 2
    if nms_workspace is not None:
        check(lib.ppy_yolo_decode_hist(ptr(head), ld, n, size, an, num_classes,
                                       anchors.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), int(stride), float(scale_x_y),
                                       ptr(im_size), 1 if clip_bbox else 0, 1 if iou_aware else 0, float(factor), ptr(boxes),
                                       ptr(scores), box_offset, total_boxes, float(hist_threshold), ptr(nms_workspace),
                                       stream_ptr()), 'yolo_decode_hist')
        return
    check(lib.ppy_yolo_decode(ptr(head), ld, n, size, an, num_classes,
                              anchors.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), int(stride), float(scale_x_y),
                              ptr(im_size), 1 if clip_bbox else 0, 1 if iou_aware else 0, float(factor), ptr(boxes),
                              ptr(scores), box_offset, total_boxes, stream_ptr()), 'yolo_decode')


def yolo_box(conv_output, anchors, stride, num_classes, scale_x_y, im_size, clip_bbox, iou_aware=False,
             iou_aware_factor=0.0):
    """(get_iou_aware_score +) yolo_box, reference model/head.py:21-141. NCHW head output in."""
    xh = to_nhwc(conv_output, PPY_F32, conv_output.shape[1])
    n, size, size_w, ld = xh.shape
    if size != size_w:
        raise ValueError('yolo_box assumes square feature maps, like the reference (head.py:25-27)')
    an = len(np.asarray(anchors).reshape(-1)) // 2
    total = size * size * an
    boxes = torch.empty((n, total, 4), dtype=torch.float32, device=xh.device)
    scores = torch.empty((n, total, num_classes), dtype=torch.float32, device=xh.device)
    ims = _cuda(im_size, 'im_size').detach().float().contiguous()
    yolo_decode_nhwc(xh, ld, n, size, anchors, stride, num_classes, scale_x_y, ims, clip_bbox, iou_aware,
                     iou_aware_factor, boxes, scores, 0, total)
    return boxes, scores


def pairwise_iou(a, b):
    a = _cuda(a).detach().float().contiguous()
    b = _cuda(b).detach().float().contiguous()
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    if out.numel():
        check(lib.ppy_pairwise_iou(ptr(a), a.shape[0], ptr(b), b.shape[0], ptr(out), stream_ptr()), 'pairwise_iou')
    return out


_NMS_WS = {}


def nms_workspace(n, num_boxes, num_classes, device):
    need = ctypes.c_size_t(0)
    check(lib.ppy_matrix_nms_workspace_bytes(n, num_boxes, num_classes, ctypes.byref(need)), 'nms_workspace_bytes')
    key = (device, need.value)
    ws = _NMS_WS.get(key)
    if ws is None:
        ws = torch.empty(need.value, dtype=torch.uint8, device=device)
        _NMS_WS[key] = ws
    return ws


def matrix_nms_launch(boxes, scores, out, counts, workspace, score_threshold, post_threshold, nms_top_k, keep_top_k,
                      use_gaussian, gaussian_sigma, have_hist=False):
    n, nb, nc = scores.shape
    fn = lib.ppy_matrix_nms_batched_hist if have_hist else lib.ppy_matrix_nms_batched
    check(fn(ptr(boxes), ptr(scores), n, nb, nc, float(score_threshold), float(post_threshold),
                                     int(nms_top_k), int(keep_top_k), 1 if use_gaussian else 0, float(gaussian_sigma),
                                     ptr(out), ptr(counts), ptr(workspace), workspace.numel(), stream_ptr()),
          'matrix_nms_batched')


# ---- sparse post-processing (whole-network path): candidates straight from the decode kernel, no dense score tensor
def nms_candidate_workspace(n, cap, device):
    need = ctypes.c_size_t(0)
    check(lib.ppy_nms_candidate_workspace_bytes(n, cap, ctypes.byref(need)), 'nms_candidate_workspace_bytes')
    return torch.empty(need.value, dtype=torch.uint8, device=device)


def nms_candidates_reset(workspace, n, cap):
    check(lib.ppy_nms_candidates_reset(ptr(workspace), n, cap, stream_ptr()), 'nms_candidates_reset')


def yolo_decode_candidates_nhwc(head, ld, n, size, anchors, stride, num_classes, scale_x_y, im_size, clip_bbox, iou_aware,
                                factor, boxes, box_offset, total_boxes, score_threshold, workspace, cap):
    anchors = np.ascontiguousarray(np.asarray(anchors, dtype=np.float32).reshape(-1))
    an = anchors.size // 2
    check(lib.ppy_yolo_decode_candidates(ptr(head), ld, n, size, an, num_classes,
                                         anchors.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), int(stride),
                                         float(scale_x_y), ptr(im_size), 1 if clip_bbox else 0, 1 if iou_aware else 0,
                                         float(factor), ptr(boxes), box_offset, total_boxes, float(score_threshold),
                                         ptr(workspace), cap, stream_ptr()), 'yolo_decode_candidates')


def matrix_nms_candidates_launch(boxes, num_classes, out, counts, workspace, cap, score_threshold, post_threshold, nms_top_k,
                                 keep_top_k, use_gaussian, gaussian_sigma):
    n, nb, _ = boxes.shape
    check(lib.ppy_matrix_nms_candidates(ptr(boxes), n, nb, num_classes, float(score_threshold), float(post_threshold),
                                        int(nms_top_k), int(keep_top_k), 1 if use_gaussian else 0, float(gaussian_sigma),
                                        ptr(out), ptr(counts), ptr(workspace), cap, stream_ptr()), 'matrix_nms_candidates')


def split_predictions(out, counts_host):
    """[N,keep,6] + per-image counts -> list of [M,6] tensors with the reference's -1 sentinel row."""
    preds = []
    for i, c in enumerate(counts_host):
        if c < 0:
            raise _lib.KernelError('matrix_nms: candidate overflow on image %d (code %d): more than 8192 scores tie '
                                   'at the top-k cutoff, nms_top_k<=0 with more than 4000 candidates, or (sparse '
                                   'post-processing) more scores above score_threshold than the candidate list holds -- '
                                   "set model.postprocess_impl = 'dense'" % (i, c))
        preds.append(out[i, :c] if c > 0 else torch.full((1, 6), -1.0, device=out.device))
    return preds


def matrix_nms_batched(boxes, scores, score_threshold, post_threshold, nms_top_k, keep_top_k, use_gaussian=False,
                       gaussian_sigma=2.):
    """Reference model/matrix_nms.py:102-151 for every image of the batch; returns a list of [M,6] tensors."""
    boxes = _cuda(boxes, 'boxes').detach().float().contiguous()
    scores = _cuda(scores, 'scores').detach().float().contiguous()
    n, nb, nc = scores.shape
    out = torch.empty((n, keep_top_k, 6), dtype=torch.float32, device=boxes.device)
    counts = torch.empty((n,), dtype=torch.int32, device=boxes.device)
    ws = nms_workspace(n, nb, nc, boxes.device)
    matrix_nms_launch(boxes, scores, out, counts, ws, score_threshold, post_threshold, nms_top_k, keep_top_k,
                      use_gaussian, gaussian_sigma)
    return split_predictions(out, counts.cpu().tolist())     # the single D2H sync of the post-processing
