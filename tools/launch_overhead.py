"""GPU-side fixed cost of one conv launch inside a CUDA graph: a graph of 64 back-to-back launches of a one-tile conv
(and of mid-size shapes), replayed; reports us per launch.  usage: python tools/launch_overhead.py"""
import os, sys
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, os.path.join(REPO, 'pytorch-ppyolo_b200')); sys.path.insert(0, REPO)
import torch
from ppyolo_b200 import ops as o
from ppyolo_b200._lib import PPY_BF16
dev = torch.device('cuda')
for (n, h, cin, cout, k) in [(1, 8, 64, 64, 1), (32, 19, 64, 256, 1), (32, 19, 1024, 512, 1), (32, 19, 2048, 512, 1), (32, 38, 256, 256, 3)]:
    w = torch.randn(cout, cin, k, k, device=dev) * 0.05
    packed = o.pack_weight(w, PPY_BF16)
    x = torch.randn(n, h, h, cin, device=dev).to(torch.bfloat16)
    y = torch.empty(n, h, h, cout, device=dev, dtype=torch.bfloat16)
    scale, shift = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
    run = lambda: o.conv_nhwc(x, packed, cin, cout, k, 1, (k - 1) // 2, scale, shift, 1, PPY_BF16, out=y)
    run(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        run()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    L = 64
    with torch.cuda.graph(g, stream=s):
        for _ in range(L):
            run()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    print('n=%d h=%d cin=%d cout=%d k=%d: %.2f us per launch in-graph' % (n, h, cin, cout, k, e0.elapsed_time(e1) / 10 / L * 1e3))
