#!/usr/bin/env python
"""DCNv2 training layer in isolation (stage-5 shape of ppyolo_2x: 512 -> 512, 19x19): forward + backward of
conv_autograd.dcnv2_kernels through the C ABI, CUDA-event time per call; `--once` runs one forward+backward between
cudaProfilerStart/Stop (for an ncu capture of dcn_backward_sample_kernel and its neighbours)."""
import argparse, json, os, sys
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, os.path.join(REPO, 'pytorch-ppyolo_b200')); sys.path.insert(0, REPO)
import torch
from ppyolo_b200.conv_autograd import dcnv2_kernels

ap = argparse.ArgumentParser()
ap.add_argument('--once', action='store_true'); ap.add_argument('--batch', type=int, default=8)
ap.add_argument('--channels', type=int, default=512); ap.add_argument('--hw', type=int, default=19); ap.add_argument('--stride', type=int, default=1)
a = ap.parse_args()
dev = torch.device('cuda')
g = torch.Generator().manual_seed(0)
c = a.channels
x = torch.randn((a.batch, c, a.hw, a.hw), generator=g).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
ow = (torch.randn((27, c, 3, 3), generator=g) * 0.02).to(dev).requires_grad_(True)
ob = (torch.randn(27, generator=g) * 1.0).to(dev).requires_grad_(True)
w = (torch.randn((c, c, 3, 3), generator=g) * (1.0 / (9 * c) ** 0.5)).to(dev).requires_grad_(True)


def step():
    y = dcnv2_kernels(x, ow, ob, w, stride=a.stride, padding=1)
    y.backward(torch.ones_like(y))
    x.grad = ow.grad = ob.grad = w.grad = None


for _ in range(3):
    step()
torch.cuda.synchronize()
if a.once:
    torch.cuda.cudart().cudaProfilerStart(); step(); torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStop()
    sys.exit(0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    step()
e1.record(); torch.cuda.synchronize()
m = a.batch * ((a.hw + 2 - 3) // a.stride + 1) ** 2
print(json.dumps({'shape': 'bs %d, %d -> %d, %dx%d, stride %d (M = %d)' % (a.batch, c, c, a.hw, a.hw, a.stride, m),
                  'fwd_bwd_ms_eager': e0.elapsed_time(e1) / 20,
                  'flops_fwd_bwd': 3 * 2 * m * 9 * c * c, 'note': 'eager launches (about 30 kernels); gemm flops = forward + dW + dcol'}))
