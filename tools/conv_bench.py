"""Times single bf16 conv launches of given shapes (CUDA events, rotating buffers so inputs do not stay hot by accident).
usage: python tools/conv_bench.py "n,h,cin,cout,k,stride[,res]" ...     (square maps; res=1 adds a residual + relu)"""
import os, sys
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, os.path.join(REPO, 'pytorch-ppyolo_b200')); sys.path.insert(0, REPO)
import torch
from ppyolo_b200 import ops as o
from ppyolo_b200._lib import PPY_BF16

dev = torch.device('cuda')
peaks = {'tf': 1394.4, 'gbs': 6538.9}
for spec in sys.argv[1:]:
    v = [int(t) for t in spec.split(',')]
    n, h, cin, cout, k, stride = v[:6]
    res = len(v) > 6 and v[6]
    w = torch.randn(cout, cin, k, k, device=dev) * 0.05
    packed = o.pack_weight(w, PPY_BF16)
    ho = (h + 2 * ((k - 1) // 2) - k) // stride + 1
    nb = 3
    xs = [torch.randn(n, h, h, cin, device=dev).to(torch.bfloat16) for _ in range(nb)]
    ys = [torch.empty(n, ho, ho, cout, device=dev, dtype=torch.bfloat16) for _ in range(nb)]
    rs = [torch.randn(n, ho, ho, cout, device=dev).to(torch.bfloat16) for _ in range(nb)] if res else [None] * nb
    scale, shift = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)

    def run(i):
        o.conv_nhwc(xs[i % nb], packed, cin, cout, k, stride, (k - 1) // 2, scale, shift, 1, PPY_BF16, residual=rs[i % nb],
                    out=ys[i % nb])
    for i in range(5):
        run(i)
    torch.cuda.synchronize()
    iters = 30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    m = n * ho * ho
    flops = 2.0 * m * cout * cin * k * k
    byts = 2.0 * (n * h * h * cin + m * cout * (2 if res else 1) + cout * cin * k * k)
    print('%-28s %8.1f us  %7.1f TF/s (%4.1f%% of %g)  %7.1f GB/s (%4.1f%% of %g)' % (
        spec, ms * 1e3, flops / ms / 1e9, 100 * flops / ms / 1e9 / peaks['tf'], peaks['tf'], byts / ms / 1e6,
        100 * byts / ms / 1e6 / peaks['gbs'], peaks['gbs']))
