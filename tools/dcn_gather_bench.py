"""Micro-benchmark of ppy_dcn_gather at the ppyolo_2x stage-5 shape (bs 32, 19x19, C=512, 3x3): mean launch time over
back-to-back launches (operands L2-resident, as in the step) and a checksum of the output bytes (to compare kernel
variants bit for bit).  Optional args: n h w c."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'pytorch-ppyolo_b200'))
import torch
from ppyolo_b200 import ops
from ppyolo_b200._lib import PPY_BF16, lib, check

n, h, w, c, k, stride = 32, 19, 19, 512, 3, 1
if len(sys.argv) > 1:
    n, h, w, c = (int(v) for v in sys.argv[1:5])
g = torch.Generator().manual_seed(0)
x = torch.randn((n, h, w, c), generator=g).to(torch.bfloat16).cuda()
om = (torch.randn((n, h, w, 32), generator=g) * 1.5).cuda()
out = torch.empty((n * h * w, k * k * c), dtype=torch.bfloat16, device='cuda')
args = (ops.ptr(x), c, n, h, w, c, ops.ptr(om), 32, k, stride, 1, ops.ptr(out), PPY_BF16)
for _ in range(5):
    check(lib.ppy_dcn_gather(*args, ops.stream_ptr()), 'gather')
torch.cuda.synchronize()
iters = 200
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    check(lib.ppy_dcn_gather(*args, ops.stream_ptr()), 'gather')
e1.record()
torch.cuda.synchronize()
print('n=%d %dx%d c=%d  %.2f us/launch  sha1=%s' % (n, h, w, c,
      e0.elapsed_time(e1) * 1e3 / iters, hashlib.sha1(out.view(torch.int16).cpu().numpy().tobytes()).hexdigest()[:12]), flush=True)
