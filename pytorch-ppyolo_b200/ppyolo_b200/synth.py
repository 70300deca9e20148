"""Deterministic synthetic weights and inputs shared by tests, smoke() and bench.py.

Every tensor is drawn from its own ``torch.Generator`` seeded by (seed, crc32(key)), so the values do
not depend on module registration order: the same call fills this repo's modules and the reference's
modules (used by ``tests/golden/make_golden.py``) with identical numbers.
"""
import math
import zlib

import torch


def _gen(seed, key):
    g = torch.Generator(device='cpu')
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(key.encode())) % (2 ** 31 - 1))
    return g


def randomize_(module, seed=0, offset_scale=0.03, head_obj_bias=-4.0, head_cls_bias=-2.0, iou_aware=True,
               num_classes=80):
    """Fill all parameters/buffers of a PPYOLO-shaped module in place.

    conv weights: He-normal; BN: weight~U(.5,1.5), bias~N(0,.1), mean~N(0,.1), var~U(.5,1.5) so that BN
    folding is exercised; DCN ``conv_offset`` (zero-init in the reference, custom_layers.py:510-511) gets
    N(0, offset_scale) weights and N(0,1) bias so sampling positions are really deformed; the output
    convs get objectness/class bias priors so that a realistic few percent of scores pass 0.01.
    """
    sd = module.state_dict()
    # last conv of each residual branch gets a small BN gamma, otherwise the variance doubles per block
    # and random-weight activations overflow exp() in the box decode after 16 blocks
    bottleneck = any('stage2_0.conv4.' in k for k in sd)
    last = '.conv3.bn.weight' if bottleneck else '.conv2.bn.weight'
    with torch.no_grad():
        for key in sorted(sd.keys()):
            t = sd[key]
            g = _gen(seed, key)
            if key.endswith('num_batches_tracked'):
                t.zero_()
            elif 'backbone.stage' in key and key.endswith(last):
                t.copy_((torch.rand(t.shape, generator=g) + 0.5) * 0.25)
            elif key.endswith('bn.weight'):
                # E[gamma^2] * E[1/var] ~= 1 so the activation scale neither explodes nor vanishes with depth
                t.copy_((torch.rand(t.shape, generator=g) + 0.5) * 0.9)
            elif key.endswith('running_var'):
                t.copy_(torch.rand(t.shape, generator=g) + 0.5)
            elif key.endswith('running_mean') or key.endswith('bn.bias'):
                t.copy_(torch.randn(t.shape, generator=g) * 0.1)
            elif key.endswith('conv_offset.weight'):
                t.copy_(torch.randn(t.shape, generator=g) * offset_scale)
            elif key.endswith('conv_offset.bias'):
                t.copy_(torch.randn(t.shape, generator=g))
            elif 'yolo_output_convs' in key and key.endswith('conv.bias'):
                b = torch.randn(t.shape, generator=g) * 0.3
                an = 3
                per = t.numel() // an
                start = an if per == num_classes + 6 else 0
                stride = num_classes + 5
                for a in range(an):
                    b[start + a * stride + 4] += head_obj_bias
                    b[start + a * stride + 5: start + (a + 1) * stride] += head_cls_bias
                t.copy_(b)
            elif 'yolo_output_convs' in key and key.endswith('conv.weight'):
                fan_in = t.shape[1] * t.shape[2] * t.shape[3]
                t.copy_(torch.randn(t.shape, generator=g) * (1.0 / math.sqrt(fan_in)))
            elif t.dim() == 4:
                fan_in = t.shape[1] * t.shape[2] * t.shape[3]
                t.copy_(torch.randn(t.shape, generator=g) * math.sqrt(2.0 / fan_in))
            else:
                t.copy_(torch.randn(t.shape, generator=g) * 0.1)
    return module


def images(batch, size, seed=1):
    """``torch.randn(N,3,S,S)``: the real pipeline yields ~N(0,1) after mean/std normalisation."""
    return torch.randn((batch, 3, size, size), generator=_gen(seed, 'images'))


def images_u8(batch, size, seed=1):
    """Resized uint8 RGB batches [N,S,S,3] (what ``Decode.process_image_u8`` produces)."""
    return torch.randint(0, 256, (batch, size, size, 3), generator=_gen(seed, 'images_u8'), dtype=torch.uint8)


def im_sizes(batch, h=480, w=640):
    return torch.tensor([[float(h), float(w)]] * batch, dtype=torch.float32)


def nms_inputs(num_boxes, num_classes=80, seed=0, extent=608.0):
    """Config C5: boxes with cx,cy~U(0,extent), w,h~U(4,204); scores
    sigmoid(N(-4,2))*sigmoid(N(-3,2)) per class -> ~15% of entries > 0.01."""
    g = _gen(seed, 'nms')
    cxy = torch.rand((num_boxes, 2), generator=g) * extent
    wh = torch.rand((num_boxes, 2), generator=g) * 200.0 + 4.0
    boxes = torch.cat([cxy - wh / 2, cxy + wh / 2], dim=1)
    obj = torch.sigmoid(torch.randn((num_boxes, 1), generator=g) * 2.0 - 4.0)
    cls = torch.sigmoid(torch.randn((num_boxes, num_classes), generator=g) * 2.0 - 3.0)
    return boxes.contiguous(), (obj * cls).contiguous()
