"""Differentiable evaluation of the YOLOv3 head for the training step.

The head is the only trainable part under the reference's default ``freeze_at=5``; its backward comes from torch
autograd over plain tensor ops here (conv/BN/pool kernels of ATen on the GPU), NOT from this repo's CUDA kernels --
hand-written dgrad/wgrad kernels are the next row of the scope table.  ATen's convolutions run at
torch's default precision (TF32 tensor-core math on this GPU; forcing strict fp32 makes cuDNN JIT-compile a kernel per
shape, ~10 s each), so losses agree with the fp32 CPU reference to ~1e-3 rather than 1e-5.  Semantics follow the reference:
DetectionBlock.__call__ model/head.py:223-231, _get_outputs :381-398, CoordConv / SPP / DropBlock
model/custom_layers.py:256-342, BatchNorm in whatever mode the module is in (train: batch statistics)."""
import torch
import torch.nn.functional as F

from model.custom_layers import Conv2dUnit, CoordConv, SPP, DropBlock


def coord_concat(x):
    b, _, h, w = x.shape
    xs = torch.arange(w, dtype=torch.float32, device=x.device) / (w - 1) * 2.0 - 1
    ys = torch.arange(h, dtype=torch.float32, device=x.device) / (h - 1) * 2.0 - 1
    return torch.cat([x, xs.view(1, 1, 1, w).expand(b, 1, h, w), ys.view(1, 1, h, 1).expand(b, 1, h, w)], dim=1)


def drop_block(x, block_size, keep_prob):
    """Reference custom_layers.py:303-342: Bernoulli(gamma) seeds grown by a max-pool, output renormalised."""
    h = x.shape[2]
    gamma = (1.0 - keep_prob) * h * h / float(block_size * block_size * (h - block_size + 1) ** 2)
    seeds = (torch.rand(x.shape, device=x.device) < gamma).float()
    mask = 1.0 - F.max_pool2d(seeds, (block_size, block_size), stride=1, padding=1)
    return x * mask * float(mask.numel()) / mask.sum()


def conv_unit(u, x):
    if not isinstance(u.conv, torch.nn.Conv2d):
        raise NotImplementedError('DCNv2 inside the trainable head is not part of any PP-YOLO config')
    y = F.conv2d(x, u.conv.weight, u.conv.bias, stride=u.stride, padding=u.padding)
    if u.bn is not None:
        bn = u.bn
        y = F.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, bn.training,
                         0.1 if bn.momentum is None else bn.momentum, bn.eps)
        if bn.training:
            bn.num_batches_tracked += 1
    if u.act_name == 'relu':
        y = F.relu(y)
    elif u.act_name == 'leaky':
        y = F.leaky_relu(y, 0.1)
    elif u.act_name == 'mish':
        y = y * torch.tanh(F.softplus(y))
    return y


def _run_layers(layers, x):
    for ly in layers:
        if isinstance(ly, CoordConv):
            x = coord_concat(x) if ly.coord_conv else x
        elif isinstance(ly, Conv2dUnit):
            x = conv_unit(ly, x)
        elif isinstance(ly, SPP):
            x = torch.cat([x] + [F.max_pool2d(x, k, 1, k // 2) for k in (5, 9, 13)], dim=1)
        elif isinstance(ly, DropBlock):
            x = x if ly.is_test else drop_block(x, ly.block_size, ly.keep_prob)
        else:
            raise TypeError(type(ly))
    return x


def head_outputs(head, body_feats):
    n_out = len(head.anchor_masks)
    feats = body_feats[-1:-n_out - 1:-1]
    outputs, route = [], None
    for i, feat in enumerate(feats):
        if i > 0:
            feat = torch.cat([route, feat], dim=1)
        blk = head.detection_blocks[i]
        route = _run_layers(blk.layers, feat)
        tip = _run_layers(blk.tip_layers, route)
        outputs.append(conv_unit(head.yolo_output_convs[i], tip))
        if i < n_out - 1:
            route = F.interpolate(conv_unit(head.upsample_layers[2 * i], route), scale_factor=2, mode='nearest')
    return outputs
