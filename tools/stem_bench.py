"""Micro-benchmark of ppy_stem_conv3x3s2 (NCHW fp32 image -> conv 3x3/s2 3->32 + BN + act -> NHWC bf16) at the headline
shape (bs 32, 608x608): mean launch time over back-to-back launches with the 142 MB input larger than L2.
PPY_NO_STEM_UMMA=1 routes the bf16 output through the fp32 SIMT kernel.  Optional args: n hw."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'pytorch-ppyolo_b200'))
import numpy as np
import torch
from ppyolo_b200 import ops
from ppyolo_b200._lib import PPY_BF16, lib, check

n, hw = 32, 608
if len(sys.argv) > 2:
    n, hw = int(sys.argv[1]), int(sys.argv[2])
g = torch.Generator().manual_seed(0)
x = torch.randn((n, 3, hw, hw), generator=g).cuda()
w = np.ascontiguousarray((torch.randn((32, 3, 3, 3), generator=g) * 0.2).numpy())
sc = np.ascontiguousarray((torch.rand(32, generator=g) + 0.5).numpy())
sh = np.ascontiguousarray((torch.randn(32, generator=g) * 0.1).numpy())
ho = (hw - 1) // 2 + 1
y = torch.empty((n, ho, ho, 32), dtype=torch.bfloat16, device='cuda')
fp = ctypes.POINTER(ctypes.c_float)
args = (ops.ptr(x), n, hw, hw, w.ctypes.data_as(fp), sc.ctypes.data_as(fp), sh.ctypes.data_as(fp), 32, 1, ctypes.c_void_p(y.data_ptr()), 32,
        PPY_BF16)
for _ in range(5):
    check(lib.ppy_stem_conv3x3s2(*args, ops.stream_ptr()), 'stem')
torch.cuda.synchronize()
iters = 50
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    check(lib.ppy_stem_conv3x3s2(*args, ops.stream_ptr()), 'stem')
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / iters
mb = (x.numel() * 4 + y.numel() * 2) / 1e6
print('simt=%s  n=%d %dx%d  %.1f us/launch  %.2f TB/s of %.0f MB  checksum=%.6f' % (os.environ.get('PPY_NO_STEM_UMMA', '0'), n, hw, hw, us,
      mb / us, mb, y.float().abs().mean().item()), flush=True)
