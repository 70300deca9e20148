"""Differentiable evaluation of the YOLOv3 head for the training step.

The head is the only trainable part under the reference's default ``freeze_at=5``.  Two implementations of its convolutions
(``impl``):

* ``'kernels'`` (default with a bf16 backbone): every conv runs forward, input-gradient and weight-gradient on this repo's
  tcgen05 conv kernel through ``conv_autograd.conv2d_kernels`` (bf16 NHWC activations and gradients, fp32 accumulation,
  fp32 weight gradients); CoordConv's two channels enter as a separate tiny ATen conv over the constant coordinate image
  (same weight tensor, so autograd sums both gradient parts).  BatchNorm, activations, pooling, upsampling and the losses
  are ATen tensor code.
* ``'aten'`` (fp32 parity mode): plain tensor ops, backward from torch autograd over ATen/cuDNN kernels at torch's default
  precision (TF32 tensor-core math on this GPU; forcing strict fp32 makes cuDNN JIT-compile a kernel per shape, ~10 s each),
  so losses agree with the fp32 CPU reference to ~1e-3.

Semantics follow the reference:
DetectionBlock.__call__ model/head.py:223-231, _get_outputs :381-398, CoordConv / SPP / DropBlock
model/custom_layers.py:256-342, BatchNorm in whatever mode the module is in (train: batch statistics)."""
import os

import torch
import torch.nn.functional as F

from model.custom_layers import Conv2dUnit, CoordConv, SPP, DropBlock


def coord_concat(x):
    b, _, h, w = x.shape
    xs = torch.arange(w, dtype=torch.float32, device=x.device) / (w - 1) * 2.0 - 1
    ys = torch.arange(h, dtype=torch.float32, device=x.device) / (h - 1) * 2.0 - 1
    return torch.cat([x, xs.view(1, 1, 1, w).expand(b, 1, h, w), ys.view(1, 1, h, 1).expand(b, 1, h, w)], dim=1)


def drop_block(x, block_size, keep_prob):
    """Reference custom_layers.py:303-342: Bernoulli(gamma) seeds grown by a max-pool, output renormalised -- the library's
    DropBlock kernels (device-side Philox stream, no host round trip, graph-capturable); differentiable."""
    from . import ops
    return ops.drop_block(x, block_size, keep_prob)


# Activations between the head's layers on the 'kernels' path: bf16 NHWC (zero-copy between layers).  Measured on the seeded
# random net (11 batch-statistic BN layers, bs 2 x 128^2): gradient cosine against the ATen/TF32 head 0.99 next to the outputs,
# 0.91 at the deepest layer -- the same structure with exact fp32 convs gives 0.993-0.9998, with torch's own bf16 convs
# 0.85-0.88, and keeping the BN/activation chain in fp32 (True) does not move it: the noise is that of bf16 GEMM operands.
ACT_FP32 = False


_COORD_IMAGES = {}


def _coord_image(h, w, device):
    """[1, 2, h, w] CoordConv channels (x then y in [-1, 1]); constant per map size, so built once (a dozen tiny arange / mul / add
    launches per use otherwise -- ~100 graph nodes of the training step)."""
    key = (h, w, str(device))
    img = _COORD_IMAGES.get(key)
    if img is None:
        xs = torch.arange(w, dtype=torch.float32, device=device) / (w - 1) * 2.0 - 1
        ys = torch.arange(h, dtype=torch.float32, device=device) / (h - 1) * 2.0 - 1
        img = torch.stack([xs.view(1, w).expand(h, w), ys.view(h, 1).expand(h, w)]).unsqueeze(0).contiguous()
        _COORD_IMAGES[key] = img
    return img


def conv_unit(u, x, impl='aten', coord=False):
    """One Conv2dUnit.  ``coord`` (kernels only): the unit's weight has two extra CoordConv input channels that ``x`` lacks."""
    if not isinstance(u.conv, torch.nn.Conv2d):           # DCNv2 unit (stage 5 of an unfrozen ResNet50-vd)
        d = u.conv
        if d.dcn_bias is not None or coord:
            raise NotImplementedError('DCNv2 with a bias / behind a CoordConv is not part of any PP-YOLO config')
        if impl == 'kernels':
            from .conv_autograd import dcnv2_kernels
            y = dcnv2_kernels(x, d.conv_offset.weight, d.conv_offset.bias, d.dcn_weight, stride=d.stride, padding=d.padding)
        else:
            raise NotImplementedError("DCNv2 has no ATen path: use the 'kernels' implementation (train_precision='bf16')")
    elif impl == 'kernels':
        from .conv_autograd import conv2d_kernels
        w = u.conv.weight
        c_main = w.shape[1] - (2 if coord else 0)
        fold = coord and u.stride == 1 and c_main % 8 == 0           # CoordConv inside the conv function (bias map + extra wgrad columns)
        y = conv2d_kernels(x, w, u.conv.bias, padding=u.padding, c_main=c_main, out_f32=ACT_FP32 or u.bn is None, stride=u.stride,
                           coord=fold)
        if coord and not fold:
            y = y + F.conv2d(_coord_image(x.shape[2], x.shape[3], x.device), w[:, c_main:], None, u.stride, u.padding).to(y.dtype)
    else:
        y = F.conv2d(x, u.conv.weight, u.conv.bias, stride=u.stride, padding=u.padding)
    if u.bn is not None:
        bn = u.bn
        momentum = 0.1 if bn.momentum is None else bn.momentum
        if bn.training:
            _PENDING_BN.append(bn)                  # counters advance in one multi-tensor launch (flush_bn_counters)
        if (impl == 'kernels' and BN_KERNELS and bn.training and y.is_cuda and y.dtype in (torch.bfloat16, torch.float32)
                and u.act_name in (None, 'relu', 'leaky') and y.shape[1] % 8 == 0 and y.shape[1] <= 2048):
            return _BnActFn.apply(y, bn.weight, bn.bias, bn.running_mean, bn.running_var, momentum, bn.eps,
                                  {None: 0, 'relu': 1, 'leaky': 2}[u.act_name])
        y = F.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, bn.training, momentum, bn.eps)
    if u.act_name == 'relu':
        y = F.relu(y)
    elif u.act_name == 'leaky':
        y = F.leaky_relu(y, 0.1)
    elif u.act_name == 'mish':
        y = y * torch.tanh(F.softplus(y))
    return y


class _SppFn(torch.autograd.Function):
    """SPP (reference model/custom_layers.py:275-290) on the library's kernels for the training head: forward = ppy_spp (three
    cascaded 5x5 max-pools + concat in one launch), backward = ppy_spp_backward (arg-max routing in shared memory, torch's
    first-maximum rule) -- ATen's NHWC max-pool kernels took 1.2 ms per step for these 19x19 maps."""

    @staticmethod
    def forward(ctx, x):
        from . import ops
        from ._lib import lib, check, PPY_BF16, PPY_F32
        n, c, h, w = x.shape
        xh = x.permute(0, 2, 3, 1)
        if not xh.is_contiguous():
            xh = xh.contiguous()
        code = PPY_BF16 if x.dtype == torch.bfloat16 else PPY_F32
        y = torch.empty((n, h, w, 4 * c), dtype=x.dtype, device=x.device)
        check(lib.ppy_spp(ops.ptr(xh), c, ops.ptr(y), 4 * c, n, h, w, c, code, ops.stream_ptr()), 'spp')
        ctx.save_for_backward(xh)
        ctx.code = code
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        from . import ops
        from ._lib import lib, check
        xh, = ctx.saved_tensors
        n, h, w, c = xh.shape
        dyh = dy.permute(0, 2, 3, 1)
        if not dyh.is_contiguous() or dyh.dtype != xh.dtype:
            dyh = dyh.to(xh.dtype).contiguous()
        dx = torch.empty_like(xh)
        check(lib.ppy_spp_backward(ops.ptr(xh), c, ops.ptr(dyh), 4 * c, ops.ptr(dx), c, n, h, w, c, ctx.code, ops.stream_ptr()), 'spp_backward')
        return dx.permute(0, 3, 1, 2)


class _BnActFn(torch.autograd.Function):
    """Train-mode BatchNorm2d + activation of a head Conv2dUnit: forward = ONE cooperative launch (ppy_bn_train_fused: batch
    statistics, running-stat update, normalise + relu / leaky) instead of ATen's statistics + transform + activation kernels;
    backward = ATen's activation and batch-norm backward on the saved batch mean / inverse std."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps, act_code):
        from . import ops
        from ._lib import lib, check, PPY_BF16, PPY_F32
        n, c, h, w = x.shape
        xh = x.permute(0, 2, 3, 1)
        if not xh.is_contiguous():
            xh = xh.contiguous()
        code = PPY_BF16 if x.dtype == torch.bfloat16 else PPY_F32
        dev = x.device
        y = torch.empty_like(xh)
        scale, shift = torch.empty(c, dtype=torch.float32, device=dev), torch.empty(c, dtype=torch.float32, device=dev)
        mean, invstd = torch.empty(c, dtype=torch.float32, device=dev), torch.empty(c, dtype=torch.float32, device=dev)
        ws = _bn_workspace(c, dev)
        check(lib.ppy_bn_train_fused(ops.ptr(xh), c, ops.ptr(y), c, n * h * w, c, code, ops.ptr(weight.detach()), ops.ptr(bias.detach()),
                                     float(eps), float(momentum), ops.ptr(running_mean), ops.ptr(running_var), ops.ptr(scale), ops.ptr(shift),
                                     None, 0, act_code, ops.ptr(ws), ops.ptr(mean), ops.ptr(invstd), ops.stream_ptr()), 'bn_train_fused')
        ctx.save_for_backward(xh, y, weight, mean, invstd)
        ctx.meta = (act_code, float(eps))
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        xh, y, weight, mean, invstd = ctx.saved_tensors
        act_code, eps = ctx.meta
        if BN_BWD_KERNEL and act_code in (0, 1, 2):
            # activation backward + both BatchNorm reductions + dx in ONE cooperative launch (ppy_bn_act_backward)
            from . import ops
            from ._lib import lib, check, PPY_BF16, PPY_F32
            n, h, w, c = xh.shape
            dyh = dy.permute(0, 2, 3, 1)
            if not dyh.is_contiguous() or dyh.dtype != xh.dtype:
                dyh = dyh.to(xh.dtype).contiguous()
            dxh = torch.empty_like(xh)
            dw, db = torch.empty(c, dtype=torch.float32, device=xh.device), torch.empty(c, dtype=torch.float32, device=xh.device)
            code = PPY_BF16 if xh.dtype == torch.bfloat16 else PPY_F32
            check(lib.ppy_bn_act_backward(ops.ptr(dyh), c, ops.ptr(xh), c, ops.ptr(y), c, ops.ptr(dxh), c, n * h * w, c, code,
                                          ops.ptr(weight.detach()), ops.ptr(mean), ops.ptr(invstd), act_code, ops.ptr(dw), ops.ptr(db),
                                          ops.ptr(_bn_workspace(c, xh.device)), ops.stream_ptr()), 'bn_act_backward')
            return dxh.permute(0, 3, 1, 2), dw.to(weight.dtype), db.to(weight.dtype), None, None, None, None, None
        x = xh.permute(0, 3, 1, 2)
        g = dy
        if act_code == 1:
            g = torch.where(y.permute(0, 3, 1, 2) > 0, dy, torch.zeros((), dtype=dy.dtype, device=dy.device))
        elif act_code == 2:
            g = torch.ops.aten.leaky_relu_backward(dy, y.permute(0, 3, 1, 2), 0.1, True)
        dx, dw, db = torch.ops.aten.native_batch_norm_backward(g, x, weight, None, None, mean, invstd, True, eps, [True, True, True])
        return dx, dw, db, None, None, None, None, None


_BN_WS = {}


def _bn_workspace(c, device):
    """Self-cleaning statistics workspace of ppy_bn_train_fused (zero on entry, zero on exit): one per channel count and device
    -- launches on one stream run one after the other."""
    key = (c, str(device))
    ws = _BN_WS.get(key)
    if ws is None:
        ws = torch.zeros(32 * c + 1, dtype=torch.float64, device=device)       # PPY_BN_WORKSPACE_DOUBLES(c)
        _BN_WS[key] = ws
    return ws


BN_KERNELS = os.environ.get('PPY_HEAD_BN_KERNELS', '1') != '0'
BN_BWD_KERNEL = os.environ.get('PPY_HEAD_BN_BWD_KERNEL', '1') != '0'


def spp_kernels_ok(x):
    return x.is_cuda and x.dtype in (torch.bfloat16, torch.float32) and x.shape[1] % 32 == 0 and \
        x.shape[2] * x.shape[3] * 32 * (5 + 2 * x.element_size()) <= 200 * 1024 and x.shape[3] <= 255


_PENDING_BN = []


def flush_bn_counters():
    """``num_batches_tracked += 1`` of every train-mode BatchNorm evaluated since the last flush -- one ``_foreach_add_`` instead
    of one tiny kernel (and CUDA-graph node) per layer."""
    if _PENDING_BN:
        torch._foreach_add_([bn.num_batches_tracked for bn in _PENDING_BN], 1)
        del _PENDING_BN[:]


def _run_layers(layers, x, impl='aten'):
    pending_coord = False
    for ly in layers:
        if isinstance(ly, CoordConv):
            if impl == 'kernels':
                pending_coord = bool(ly.coord_conv)          # folded into the next conv instead of concatenated
            else:
                x = coord_concat(x) if ly.coord_conv else x
        elif isinstance(ly, Conv2dUnit):
            x = conv_unit(ly, x, impl, pending_coord)
            pending_coord = False
        elif isinstance(ly, SPP):
            if impl == 'kernels' and ly.seq == 'asc' and spp_kernels_ok(x):
                x = _SppFn.apply(x)
            else:
                x = torch.cat([x] + [F.max_pool2d(x, k, 1, k // 2) for k in (5, 9, 13)], dim=1)
        elif isinstance(ly, DropBlock):
            x = x if ly.is_test else drop_block(x, ly.block_size, ly.keep_prob)
        else:
            raise TypeError(type(ly))
    return x


def head_outputs(head, body_feats, impl='aten'):
    n_out = len(head.anchor_masks)
    feats = body_feats[-1:-n_out - 1:-1]
    if impl == 'kernels':                                     # NHWC (channels_last) activations between the layers
        feats = [f.to(torch.float32 if ACT_FP32 else torch.bfloat16).contiguous(memory_format=torch.channels_last) for f in feats]
    outputs, route = [], None
    for i, feat in enumerate(feats):
        if i > 0:
            feat = torch.cat([route, feat], dim=1)
        blk = head.detection_blocks[i]
        route = _run_layers(blk.layers, feat, impl)
        tip = _run_layers(blk.tip_layers, route, impl)
        outputs.append(conv_unit(head.yolo_output_convs[i], tip, impl).float())
        if i < n_out - 1:
            route = F.interpolate(conv_unit(head.upsample_layers[2 * i], route, impl), scale_factor=2, mode='nearest')
    flush_bn_counters()
    return outputs
