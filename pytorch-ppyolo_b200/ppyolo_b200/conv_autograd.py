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


def _dgrad(dyh, weight, c_main, stride, pad, in_hw, out_f32):
    """Input gradient of a k x k conv on the conv kernel: a stride-1 conv of dY with the 180-degree-rotated transposed weight and
    padding k-1-pad.  A strided conv's dY is first spread onto the input grid (zeros between the samples: dx = full correlation of
    the zero-stuffed dY) -- 4x redundant MACs for stride 2, three layers of an unfrozen ResNet-vd."""
    cout, _, k, _ = weight.shape
    dev = dyh.device
    o_pad = dyh.shape[-1]
    h, w = in_hw
    if stride > 1:
        n, ho, wo, _ = dyh.shape
        hz, wz = h + 2 * pad - k + 1, w + 2 * pad - k + 1          # stride-1 output grid of the same conv
        z = torch.zeros((n, hz, wz, o_pad), dtype=dyh.dtype, device=dev)
        z[:, 0:ho * stride:stride, 0:wo * stride:stride] = dyh
        dyh = z
    # rotated + transposed + packed in one launch (was: flip, transposing copy, pack)
    packed_t = ops.pack_weight_dgrad(weight, c_main, o_pad)
    dxh = ops.conv_nhwc(dyh, packed_t, o_pad, c_main, k, 1, k - 1 - pad, _const('one', c_main, dev), _const('zero', c_main, dev),
                        0, PPY_BF16, out_code=PPY_F32 if out_f32 else PPY_BF16)
    return dxh


def _kmajor(t_nhwc, c, k, stride, pad, m_pad):
    """[c*k*k rows][m_pad] K-major operand (ppy_im2col_kmajor[_strided]); columns = the conv's output pixels."""
    n, h, w, ld = t_nhwc.shape
    out = torch.empty((c * k * k, m_pad), dtype=torch.bfloat16, device=t_nhwc.device)
    check(lib.ppy_im2col_kmajor_strided(ops.ptr(t_nhwc), ld, n, h, w, c, k, stride, pad, ops.ptr(out), m_pad, ops.stream_ptr()),
          'im2col_kmajor')
    return out


def _wgrad_gemm(a_rows_kmajor, rows, b_kmajor, cols, m_pad):
    """out[rows][cols] = A[rows][m] . B[cols][m]^T on the conv kernel's partial-sum (split-K) instantiation; fp32 result."""
    dev = a_rows_kmajor.device
    a_op = a_rows_kmajor[:rows].view(1, 1, rows, m_pad)
    cols_pad = ops.round_up(cols, 8)
    out = torch.zeros((1, 1, rows, cols_pad), dtype=torch.float32, device=dev)
    ops.conv_nhwc(a_op, (b_kmajor, m_pad, m_pad, cols), m_pad, cols, 1, 1, 0, _const('one', cols, dev), _const('zero', cols, dev), 0,
                  PPY_BF16, out=out, out_code=PPY_F32, accumulate=True, split_k=0)
    return out[0, 0, :, :cols]


_COORD = {}


def coord_cols(h, w, k, pad, device):
    """im2col of CoordConv's two channels (model/custom_layers.py:256-272: x in [-1, 1] along W, then y along H; zero padding like
    any input channel) for a k x k stride-1 conv: fp32 [h*w, 2*k*k], columns in the weight's (channel, ky, kx) order.  Constant per
    map size: built once."""
    key = ('cols', h, w, k, pad, str(device))
    t = _COORD.get(key)
    if t is None:
        xs = torch.arange(w, dtype=torch.float32, device=device) / (w - 1) * 2.0 - 1
        ys = torch.arange(h, dtype=torch.float32, device=device) / (h - 1) * 2.0 - 1
        img = torch.stack([xs.view(1, w).expand(h, w), ys.view(h, 1).expand(h, w)]).unsqueeze(0)
        t = torch.nn.functional.unfold(img, k, padding=pad)[0].t().contiguous()                  # [h*w, 2*k*k]
        _COORD[key] = t
    return t


def _coord_rows_kmajor(n, h, w, k, pad, m_pad, device):
    """The same columns as the K-major bf16 operand of the weight-gradient GEMM: [8-padded 2*k*k rows][m_pad], tiled over the batch."""
    key = ('rows', n, h, w, k, pad, m_pad, str(device))
    t = _COORD.get(key)
    if t is None:
        cols = coord_cols(h, w, k, pad, device)                                                # [h*w, 2kk]
        rows = ops.round_up(cols.shape[1], 8)
        t = torch.zeros((rows, m_pad), dtype=torch.bfloat16, device=device)
        t[:cols.shape[1], :n * h * w] = cols.t().repeat(1, n).to(torch.bfloat16)
        _COORD[key] = t
    return t


def _wgrad(xh, dyh, weight, c_main, stride, pad, coord=False):
    """Weight gradient dW = dY^T [O x M] . Xcol [M x C k k] (M = the output pixels), result in the weight's OIHW order.  ``coord``:
    the weight's channels [c_main, c_main + 2) are CoordConv's -- their constant im2col columns join the GEMM's B operand, so the
    same launch yields their gradient."""
    cout, _, k, _ = weight.shape
    n, ho, wo, _ = dyh.shape
    m_pad = ops.round_up(n * ho * wo, 64)
    c_eff = c_main if c_main % 8 == 0 else xh.shape[-1]        # the stem's 3 channels travel zero-padded to 8
    kk = c_eff * k * k
    a_rows = ops.round_up(cout, 8)
    a_full = _kmajor(dyh, a_rows, 1, 1, 0, m_pad)
    if coord:
        crow = _coord_rows_kmajor(n, ho, wo, k, pad, m_pad, xh.device)
        b_op = torch.empty((kk + crow.shape[0], m_pad), dtype=torch.bfloat16, device=xh.device)
        check(lib.ppy_im2col_kmajor_strided(ops.ptr(xh), xh.shape[-1], n, xh.shape[1], xh.shape[2], c_eff, k, stride, pad, ops.ptr(b_op), m_pad,
                                            ops.stream_ptr()), 'im2col_kmajor')
        b_op[kk:].copy_(crow)
        out = _wgrad_gemm(a_full, cout, b_op, kk + 2 * k * k, m_pad)
        return out.reshape(cout, c_main + 2, k, k) if c_eff == c_main else None
    b_op = _kmajor(xh, c_eff, k, stride, pad, m_pad)
    out = _wgrad_gemm(a_full, cout, b_op, kk, m_pad)
    if c_eff == weight.shape[1]:                               # the GEMM result IS the gradient (a view when its rows are unpadded)
        return out.reshape(cout, c_eff, k, k)
    dw = torch.zeros_like(weight, dtype=torch.float32)
    dw[:, :c_main] = out.reshape(cout, c_eff, k, k)[:, :c_main]
    return dw


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, c_main, pad, out_f32, stride, coord):
        if not x.is_cuda:
            raise RuntimeError('ppyolo_b200: conv2d_kernels needs CUDA tensors -- there is no CPU fallback')
        cout, _, k, _ = weight.shape
        packed = ops.pack_weight(weight, PPY_BF16, c_begin=0, c_count=c_main, cache=False)
        xh = _nhwc_bf16(x, packed[1])
        shift = bias.detach().float().contiguous() if bias is not None else _const('zero', cout, x.device)
        bias_map = None
        if coord:
            # CoordConv's two channels (the weight's last two) contribute a batch-invariant per-pixel term: one small fp32 matmul of
            # their constant im2col columns, added to the accumulator by the conv's epilogue (`bias_map`) -- no concat, no extra pass
            wc = weight.detach()[:, c_main:c_main + 2].reshape(cout, 2 * k * k).float()
            bias_map = torch.matmul(coord_cols(x.shape[2], x.shape[3], k, pad, x.device), wc.t()).contiguous()       # [h*w, cout]
        y = ops.conv_nhwc(xh, packed, c_main, cout, k, stride, pad, _const('one', cout, x.device), shift, 0, PPY_BF16,
                          out_code=PPY_F32 if out_f32 else PPY_BF16, bias_map=bias_map)
        ctx.save_for_backward(xh, weight)
        ctx.meta = (c_main, pad, bias is not None, tuple(x.shape), x.dtype, stride, coord)
        return y[..., :cout].permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        xh, weight = ctx.saved_tensors
        c_main, pad, has_bias, x_shape, x_dtype, stride, coord = ctx.meta
        cout = weight.shape[0]
        dyh = _nhwc_bf16(dy, ops.round_up(cout, 8))
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dxh = _dgrad(dyh, weight, c_main, stride, pad, x_shape[2:], x_dtype != torch.bfloat16)
            dx = dxh[..., :c_main].permute(0, 3, 1, 2)
            if dx.dtype != x_dtype:
                dx = dx.to(x_dtype)
        if ctx.needs_input_grad[1]:
            dw = _wgrad(xh, dyh, weight, c_main, stride, pad, coord)
        if has_bias and ctx.needs_input_grad[2]:
            db = dy.float().sum(dim=(0, 2, 3))
        return dx, dw, db, None, None, None, None, None


def conv2d_kernels(x, weight, bias=None, padding=0, c_main=None, out_f32=False, stride=1, coord=False):
    """conv2d over the first ``c_main`` input channels of ``weight``.  ``coord``: the weight has exactly two more input channels,
    CoordConv's (x then y coordinate), which ``x`` does not carry -- their contribution and their weight gradient are computed
    inside (stride 1, c_main % 8 == 0); without ``coord`` extra weight channels are the caller's business and get a zero gradient.
    ``x``: logical [N, c_main, H, W]; returns logical [N, cout, Ho, Wo] (channels_last bf16, or fp32 with ``out_f32``)."""
    c_main = weight.shape[1] if c_main is None else c_main
    if x.shape[1] != c_main:
        raise ValueError('conv2d_kernels: input has %d channels, expected %d' % (x.shape[1], c_main))
    if coord and (stride != 1 or c_main % 8 != 0 or weight.shape[1] != c_main + 2):
        raise ValueError('conv2d_kernels: coord needs stride 1, c_main % 8 == 0 and exactly two extra weight channels')
    return _ConvFn.apply(x, weight, bias, c_main, padding, out_f32, stride, coord)


class _DcnFn(torch.autograd.Function):
    """DCNv2.forward (reference model/custom_layers.py:551-677) with its backward on this repo's kernels.

    forward   offset/mask conv (conv kernel, fp32 NHWC result) -> whole-layer fused deformable conv (dcn_umma.cu)
    backward  cols  = ppy_dcn_gather(x, om)                      [M x 9C] bf16 (rebuilt, not saved)
              dW    = dY^T . cols                                partial-sum GEMM, fp32
              dcol  = dY . Wt                                    1x1 conv with 9C output channels, fp32
              (dx, d_om) = ppy_dcn_backward_sample(dcol, x, om)   vector atomics into fp32 dx
              conv_offset: dgrad added into dx, wgrad, bias      the strided conv backward above
    The native spec the reference vendors without loading is external/DCNv2/src/cuda/dcn_v2_im2col_cuda.cu:197-327."""

    @staticmethod
    def forward(ctx, x, offset_w, offset_b, dcn_w, stride, pad):
        if not x.is_cuda:
            raise RuntimeError('ppyolo_b200: dcnv2_kernels needs CUDA tensors -- there is no CPU fallback')
        cout, cin, k, _ = dcn_w.shape
        dev = x.device
        n_om = offset_w.shape[0]
        om_packed = ops.pack_weight(offset_w, PPY_BF16, cache=False)
        xh = _nhwc_bf16(x, om_packed[1])
        om = ops.conv_nhwc(xh, om_packed, cin, n_om, k, stride, pad, _const('one', n_om, dev), offset_b.detach().float().contiguous(), 0,
                           PPY_BF16, out_code=PPY_F32)
        packed = ops.pack_weight(dcn_w, PPY_BF16, cache=False)
        y = ops.conv_nhwc(xh, packed, cin, cout, k, stride, pad, _const('one', cout, dev), _const('zero', cout, dev), 0, PPY_BF16,
                          offset_mask=om)
        ctx.save_for_backward(xh, om, offset_w, dcn_w)
        ctx.meta = (stride, pad, tuple(x.shape), x.dtype)
        return y[..., :cout].permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        xh, om, offset_w, dcn_w = ctx.saved_tensors
        stride, pad, x_shape, x_dtype = ctx.meta
        cout, cin, k, _ = dcn_w.shape
        n, _, h, w = x_shape
        dev = dy.device
        taps = k * k
        o_pad = ops.round_up(cout, 8)
        dyh = _nhwc_bf16(dy, o_pad)
        _, ho, wo, om_ld = om.shape
        m = n * ho * wo
        m_pad = ops.round_up(m, 64)
        kk = taps * cin
        # weight gradient over the re-sampled matrix (columns tap*C + c, the packed 3x3 weight's own K order)
        cols = torch.empty((1, 1, m, kk), dtype=torch.bfloat16, device=dev)
        check(lib.ppy_dcn_gather(ops.ptr(xh), xh.shape[-1], n, h, w, cin, ops.ptr(om), om_ld, k, stride, pad, ops.ptr(cols), PPY_BF16,
                                 ops.stream_ptr()), 'dcn_gather')
        dw = None
        if ctx.needs_input_grad[3]:
            a_full = _kmajor(dyh, o_pad, 1, 1, 0, m_pad)
            b_op = _kmajor(cols, kk, 1, 1, 0, m_pad)
            dw = _wgrad_gemm(a_full, cout, b_op, kk, m_pad).reshape(cout, k, k, cin).permute(0, 3, 1, 2).contiguous()
        del cols
        # dcol = dY . Wt: a 1x1 conv whose output channel tap*C + c carries W[:, c, tap]
        wt = torch.zeros((kk, o_pad, 1, 1), dtype=torch.float32, device=dev)
        wt[:, :cout, 0, 0] = dcn_w.detach().permute(2, 3, 1, 0).reshape(kk, cout)
        packed_t = ops.pack_weight(wt, PPY_BF16, cache=False)
        dcol = ops.conv_nhwc(dyh, packed_t, o_pad, kk, 1, 1, 0, _const('one', kk, dev), _const('zero', kk, dev), 0, PPY_BF16,
                             out_code=PPY_F32)
        d_om = torch.zeros_like(om)
        dx32 = torch.zeros((n, h, w, cin), dtype=torch.float32, device=dev)
        check(lib.ppy_dcn_backward_sample(ops.ptr(xh), xh.shape[-1], n, h, w, cin, ops.ptr(om), om_ld, k, stride, pad, ops.ptr(dcol),
                                          ops.ptr(dx32), cin, ops.ptr(d_om), PPY_BF16, ops.stream_ptr()), 'dcn_backward_sample')
        del dcol
        # the offset/mask conv's own backward
        n_om = offset_w.shape[0]
        d_omh = d_om.to(torch.bfloat16)
        dx = dow = dob = None
        if ctx.needs_input_grad[0]:
            dx32 = dx32 + _dgrad(d_omh, offset_w, cin, stride, pad, (h, w), True)[..., :cin]
            dx = dx32.permute(0, 3, 1, 2).to(x_dtype)
        if ctx.needs_input_grad[1]:
            dow = _wgrad(xh, d_omh, offset_w, cin, stride, pad)
        if ctx.needs_input_grad[2]:
            dob = d_om[..., :n_om].sum(dim=(0, 1, 2))
        return dx, dow, dob, dw, None, None


def dcnv2_kernels(x, offset_w, offset_b, dcn_w, stride=1, padding=1):
    """Differentiable DCNv2 (3x3, deformable_groups 1, no bias): logical NCHW in, channels_last bf16 out."""
    return _DcnFn.apply(x, offset_w, offset_b, dcn_w, stride, padding)
