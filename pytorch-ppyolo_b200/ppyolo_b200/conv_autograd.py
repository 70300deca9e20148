"""Differentiable stride-1 convolution whose forward, input gradient and weight gradient all run on this repo's tcgen05
conv kernel (``ppy_conv_bf16``) -- the GEMM-class work of the trainable YOLOv3 head (reference model/head.py:223-231,
:381-398 through Conv2dUnit.forward model/custom_layers.py:243-253) in the training step.

    forward   y  = conv(x, W)                       one launch, TMA-fed (tma_a / slab / patch / im2col by shape)
    dgrad     dx = conv(dy, W^T rotated 180 deg)    the SAME kernel: a k x k stride-1 conv with pad k-1-p over dy
    wgrad     dW = dY^T [O x M] . Xcol [M x C k k]  the kernel's partial-sum (split-K, fp32 atomics) instantiation as a 1x1
              "conv" whose rows are the output channels and whose K runs over the M = n*h*w pixels; both operands are
              K-major matrices built by ATen data movement (transpose of dy, unfold + transpose of x), the result is
              [O][C*k*k] in the weight's own OIHW order

The GEMM operands are bf16 NHWC, accumulation is fp32.  Outputs / input gradients leave in the dtype the caller works in:
bf16 (torch ``channels_last`` views, zero-copy between layers) or fp32 (``out_f32`` / fp32 inputs: only the GEMM operands are
rounded, the BatchNorm / activation chain in between keeps full precision); weight gradients are always fp32.  Everything else of the head (BatchNorm, activations, pooling, upsampling, the
losses) stays ATen tensor code.  No CPU / ATen fallback for the convolution itself: non-CUDA tensors raise."""
import torch

from . import ops
from ._lib import PPY_BF16, PPY_F32, lib, check

_CONST = {}


def _const(kind, n, dev):
    key = (kind, n, dev)
    t = _CONST.get(key)
    if t is None:
        t = (torch.ones if kind == 'one' else torch.zeros)(n, dtype=torch.float32, device=dev)
        _CONST[key] = t
    return t


def _nhwc_bf16(x, c_pad):
    """Logical NCHW tensor -> contiguous [N,H,W,c_pad] bf16 (zero-copy for channels_last bf16 tensors of the right width)."""
    n, c, h, w = x.shape
    xh = x.permute(0, 2, 3, 1)
    if x.dtype == torch.bfloat16 and c == c_pad and xh.is_contiguous():
        return xh
    if c == c_pad:
        return xh.to(torch.bfloat16).contiguous()
    out = torch.zeros((n, h, w, c_pad), dtype=torch.bfloat16, device=x.device)
    out[..., :c] = xh
    return out


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, c_main, pad, out_f32):
        if not x.is_cuda:
            raise RuntimeError('ppyolo_b200: conv2d_kernels needs CUDA tensors -- there is no CPU fallback')
        cout, _, k, _ = weight.shape
        packed = ops.pack_weight(weight, PPY_BF16, c_begin=0, c_count=c_main, cache=False)
        xh = _nhwc_bf16(x, packed[1])
        shift = bias.detach().float().contiguous() if bias is not None else _const('zero', cout, x.device)
        y = ops.conv_nhwc(xh, packed, c_main, cout, k, 1, pad, _const('one', cout, x.device), shift, 0, PPY_BF16,
                          out_code=PPY_F32 if out_f32 else PPY_BF16)
        ctx.save_for_backward(xh, weight)
        ctx.meta = (c_main, pad, bias is not None, tuple(x.shape), x.dtype)
        return y[..., :cout].permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        xh, weight = ctx.saved_tensors
        c_main, pad, has_bias, x_shape, x_dtype = ctx.meta
        cout, cin_total, k, _ = weight.shape
        n, _, h, w = x_shape
        dev = dy.device
        o_pad = ops.round_up(cout, 8)
        dyh = _nhwc_bf16(dy, o_pad)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            # dgrad: W^T rotated by 180 degrees, [c_main, cout(+pad), k, k]
            wt = torch.zeros((c_main, o_pad, k, k), dtype=torch.float32, device=dev)
            wt[:, :cout] = weight.detach()[:, :c_main].flip(2, 3).permute(1, 0, 2, 3)
            packed_t = ops.pack_weight(wt, PPY_BF16, cache=False)
            dxh = ops.conv_nhwc(dyh, packed_t, o_pad, c_main, k, 1, k - 1 - pad, _const('one', c_main, dev), _const('zero', c_main, dev),
                                0, PPY_BF16, out_code=PPY_BF16 if x_dtype == torch.bfloat16 else PPY_F32)
            dx = dxh[..., :c_main].permute(0, 3, 1, 2)
            if dx.dtype != x_dtype:
                dx = dx.to(x_dtype)
        if ctx.needs_input_grad[1]:
            m = n * h * w
            m_pad = ops.round_up(m, 64)
            kk = c_main * k * k
            # B operand: Xcol^T [C*k*k][M] (K-major in the pixel index), rows in the weight's (c, ky, kx) order; A operand: dY^T
            # [cout][M] -- both written in one pass each by ppy_im2col_kmajor (zero padding of M included)
            b_op = torch.empty((kk, m_pad), dtype=torch.bfloat16, device=dev)
            check(lib.ppy_im2col_kmajor(ops.ptr(xh), xh.shape[-1], n, h, w, c_main, k, pad, ops.ptr(b_op), m_pad, ops.stream_ptr()), 'im2col_kmajor')
            a_rows = ops.round_up(cout, 8)
            a_full = torch.empty((a_rows, m_pad), dtype=torch.bfloat16, device=dev)
            check(lib.ppy_im2col_kmajor(ops.ptr(dyh), dyh.shape[-1], n, h, w, a_rows, 1, 0, ops.ptr(a_full), m_pad, ops.stream_ptr()), 'transpose_kmajor')
            a_op = a_full[:cout].view(1, 1, cout, m_pad)
            kk_pad = ops.round_up(kk, 8)
            out = torch.zeros((1, 1, cout, kk_pad), dtype=torch.float32, device=dev)
            ops.conv_nhwc(a_op, (b_op, m_pad, m_pad, kk), m_pad, kk, 1, 1, 0, _const('one', kk, dev), _const('zero', kk, dev), 0,
                          PPY_BF16, out=out, out_code=PPY_F32, accumulate=True, split_k=0)
            dw = torch.zeros_like(weight, dtype=torch.float32)
            dw[:, :c_main] = out[0, 0, :, :kk].reshape(cout, c_main, k, k)
        if has_bias and ctx.needs_input_grad[2]:
            db = dy.float().sum(dim=(0, 2, 3))
        return dx, dw, db, None, None, None


def conv2d_kernels(x, weight, bias=None, padding=0, c_main=None, out_f32=False):
    """Stride-1 conv2d over the first ``c_main`` input channels of ``weight`` (the rest -- CoordConv's two -- is the caller's).
    ``x``: logical [N, c_main, H, W]; returns logical [N, cout, H, W] (channels_last bf16, or fp32 with ``out_f32``)."""
    c_main = weight.shape[1] if c_main is None else c_main
    if x.shape[1] != c_main:
        raise ValueError('conv2d_kernels: input has %d channels, expected %d' % (x.shape[1], c_main))
    return _ConvFn.apply(x, weight, bias, c_main, padding, out_f32)
