"""Differentiable evaluation of the ResNet-vd backbones for training with ``freeze_at < 5`` (reference model/resnet_vd.py:132-168,
:302-330; blocks :15-87, :224-267; the stage-5 DCNv2 units model/custom_layers.py:551-677).

The default configs freeze the whole backbone (``freeze_at=5``) and run it on the static kernel engine under ``no_grad``
(``PPYOLO.forward_train``).  With trainable stages every Conv2dUnit goes through ``autograd_head.conv_unit`` instead: convolution
forward / input gradient / weight gradient on the tcgen05 conv kernel (``conv_autograd.conv2d_kernels``, stride 2 included), the
deformable convolutions and their backward through ``conv_autograd.dcnv2_kernels`` (fused forward kernel, gather + two GEMMs +
``ppy_dcn_backward_sample``), bf16 channels_last activations between the layers.  BatchNorm (batch statistics, as the reference's
train-mode BNs -- frozen ones included), ReLU, the two pooling flavours and the residual adds are ATen tensor code,
like in the head.  The frozen prefix (stem and stages below ``freeze_at``) runs under ``no_grad`` so nothing is saved for it."""
import torch
import torch.nn.functional as F

from .autograd_head import conv_unit, flush_bn_counters, ACT_FP32


def _block(blk, x, impl):
    kind = type(blk).__name__
    if kind in ('ConvBlock', 'IdentityBlock'):
        y = conv_unit(blk.conv1, x, impl)
        y = conv_unit(blk.conv2, y, impl)
        y = conv_unit(blk.conv3, y, impl)
        if kind == 'ConvBlock':
            sc = x if blk.is_first else F.avg_pool2d(x, 2, 2, 0)
            sc = conv_unit(blk.conv4, sc, impl)
        else:
            sc = x
        return F.relu(y + sc)
    if kind == 'BasicBlock':
        y = conv_unit(blk.conv1, x, impl)
        y = conv_unit(blk.conv2, y, impl)
        if blk.conv3 is not None:
            sc = conv_unit(blk.conv3, x if blk.is_first else F.avg_pool2d(x, 2, 2, 0), impl)
        else:
            sc = x
        return F.relu(y + sc)
    raise TypeError(kind)


def _trainable(mod):
    return any(p.requires_grad for p in mod.parameters())


def backbone_features(backbone, x, impl='kernels'):
    """[C3, C4, C5] (or the configured picks) of ``backbone`` for the NCHW fp32 batch ``x``; differentiable w.r.t. every
    parameter that requires grad."""
    if impl == 'kernels':
        x = x.to(torch.float32 if ACT_FP32 else torch.bfloat16).contiguous(memory_format=torch.channels_last)
    stem = backbone._stem_units()
    grad_on = any(_trainable(u) for u in stem)
    with torch.set_grad_enabled(grad_on and torch.is_grad_enabled()):
        for u in stem:
            x = conv_unit(u, x, impl)
        x = F.max_pool2d(x, 3, 2, 1)
    outs = []
    for stage in (2, 3, 4, 5):
        for name in backbone.stage_names(stage):
            blk = getattr(backbone, name)
            grad_on = grad_on or _trainable(blk)
            with torch.set_grad_enabled(grad_on and torch.is_grad_enabled()):
                x = _block(blk, x, impl)
        if stage in backbone.feature_maps:
            outs.append(x)
    flush_bn_counters()
    return outs
