#!/usr/bin/env python
"""Micro-benchmark of one ppyolo_2x stage-5 DCNv2 layer at bs 32 (M = 11552, C = N = 512): offset conv + deformable conv,
bf16 and f16x2, CUDA events.  PPY_NO_DCN2=1 selects the old producer-mode kernel (read once per process)."""
import json, os, sys
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, os.path.join(REPO, 'pytorch-ppyolo_b200')); sys.path.insert(0, REPO)
import torch
from ppyolo_b200 import ops
from ppyolo_b200._lib import PPY_F32, PPY_BF16

dev = 'cuda'
g = torch.Generator().manual_seed(0)
res = {'no_dcn2': os.environ.get('PPY_NO_DCN2') is not None}
for stride, hw in ((1, 19), (2, 38)):
    n, c, cout = 32, 512, 512
    x = torch.randn((n, c, hw, hw), generator=g).to(dev)
    ow = (torch.randn((27, c, 3, 3), generator=g) * 0.03).to(dev)
    ob = torch.randn(27, generator=g).to(dev)
    wt = (torch.randn((cout, c, 3, 3), generator=g) / 68.0).to(dev)
    one, zero = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
    one27 = torch.ones(27, device=dev)
    xh = ops.to_nhwc(x, PPY_BF16)
    pk_o, pk_w = ops.pack_weight(ow, PPY_BF16), ops.pack_weight(wt, PPY_BF16)
    xp = ops.split_pair(ops.to_nhwc(x, PPY_F32))

    def bf16_offset():
        return ops.conv_nhwc(xh, pk_o, c, 27, 3, stride, 1, one27, ob, 0, PPY_BF16, out_code=PPY_F32)
    om = bf16_offset()

    def bf16_dcn():
        return ops.conv_nhwc(xh, pk_w, c, cout, 3, stride, 1, one, zero, 1, PPY_BF16, offset_mask=om)

    def pair_offset():
        return ops.conv_pair(xp, ow, one27, ob, stride, 1, 0, out_f32=True)
    omp = pair_offset()

    def pair_dcn():
        return ops.conv_pair(xp, wt, one, zero, stride, 1, 1, offset_mask=omp)

    for name, fn in (('bf16_offset', bf16_offset), ('bf16_dcn', bf16_dcn), ('f16x2_offset', pair_offset), ('f16x2_dcn', pair_dcn)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res['s%d_%s_us' % (stride, name)] = e0.elapsed_time(e1) / 10 * 1e3
print(json.dumps(res))
