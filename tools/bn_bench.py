#!/usr/bin/env python
"""Train-mode BatchNorm of the frozen backbone, layer shape by layer shape (config C4: bs 8 x 608^2, bf16 NHWC):
ppy_bn_train_fused (one cooperative launch: statistics, grid barrier, normalise + activation + residual) timed with CUDA events
over a CUDA graph that rotates through enough buffers to exceed the L2.  Prints microseconds and algorithmic GB/s per shape
(2 reads + 1 write of the tensor, + 1 read with a residual) and the total over the 55 backbone layers of one step.

    python tools/bn_bench.py [--batch 8] [--size 608]
"""
import argparse, ctypes, json, os, sys
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, os.path.join(REPO, 'pytorch-ppyolo_b200')); sys.path.insert(0, REPO)
import torch
from ppyolo_b200 import ops
from ppyolo_b200._lib import lib, check, PPY_BF16

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=8); ap.add_argument('--size', type=int, default=608)
ap.add_argument('--only', default=None, help='one layer shape by name (ncu captures)')
a = ap.parse_args()
dev = torch.device('cuda', 0)
S = a.size
# (name, spatial, channels, residual, layers of this shape in ResNet50-vd)
SHAPES = [('stem.conv1_1/2', S // 2, 32, False, 2), ('stem.conv1_3', S // 2, 64, False, 1),
          ('stage2.conv1/2', S // 4, 64, False, 6), ('stage2.conv3', S // 4, 256, True, 3), ('stage2.short', S // 4, 256, False, 1),
          ('stage3_0.conv1', S // 4, 128, False, 1), ('stage3.conv1/2', S // 8, 128, False, 7), ('stage3.conv3', S // 8, 512, True, 4),
          ('stage3.short', S // 8, 512, False, 1),
          ('stage4_0.conv1', S // 8, 256, False, 1), ('stage4.conv1/2', S // 16, 256, False, 11), ('stage4.conv3', S // 16, 1024, True, 6),
          ('stage4.short', S // 16, 1024, False, 1),
          ('stage5_0.conv1', S // 16, 512, False, 1), ('stage5.conv1/2', S // 32, 512, False, 5), ('stage5.conv3', S // 32, 2048, True, 3),
          ('stage5.short', S // 32, 2048, False, 1)]
total_us, total_bytes, out = 0.0, 0.0, []
for name, hw, c, has_res, count in SHAPES:
    if a.only and name != a.only:
        continue
    rows = a.batch * hw * hw
    nbytes = rows * c * 2
    nbuf = max(2, min(8, int(400e6 // nbytes) + 1))
    xs = [torch.randn(rows, c, device=dev).to(torch.bfloat16) for _ in range(nbuf)]
    ys = [torch.empty_like(x) for x in xs]
    rs = [torch.randn(rows, c, device=dev).to(torch.bfloat16) for _ in range(nbuf)] if has_res else None
    g, b = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
    sc, sh = torch.empty(c, device=dev), torch.empty(c, device=dev)
    ws = torch.zeros(32 * c + 1, dtype=torch.float64, device=dev)

    def launch(i):
        check(lib.ppy_bn_train_fused(ops.ptr(xs[i]), c, ops.ptr(ys[i]), c, rows, c, PPY_BF16, ops.ptr(g), ops.ptr(b), 1e-5, 0.1, ops.ptr(rm),
                                     ops.ptr(rv), ops.ptr(sc), ops.ptr(sh), ops.ptr(rs[i]) if has_res else None, c if has_res else 0, 1,
                                     ops.ptr(ws), None, None, ops.stream_ptr()), 'bn')
    reps = 4 * nbuf
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for i in range(nbuf):
            launch(i)
        st.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=st):
            for k in range(reps):
                launch(k % nbuf)
        graph.replay(); st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); graph.replay(); e1.record(st); st.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    # correctness of the last launch against torch (fp32 statistics of the bf16 tensor)
    x32 = xs[(reps - 1) % nbuf].float()
    m, v = x32.mean(0), x32.var(0, unbiased=False)
    want = (x32 - m) * torch.rsqrt(v + 1e-5)
    if has_res:
        want = want + rs[(reps - 1) % nbuf].float()
    err = (ys[(reps - 1) % nbuf].float() - want.clamp_min(0)).abs().max().item()
    algo = nbytes * (4 if has_res else 3)
    out.append({'layer': name, 'rows': rows, 'c': c, 'residual': has_res, 'count': count, 'us': us, 'algorithmic_GBps': algo / us * 1e-3, 'max_err': err})
    print('%-16s rows %8d c %5d res %d  x%-2d  %8.1f us  %7.0f GB/s  (tensor %6.1f MB)  err %.3g' % (name, rows, c, has_res, count, us, algo / us * 1e-3, nbytes / 1e6, err), flush=True)
    total_us += us * count; total_bytes += algo * count
    del xs, ys, rs
    torch.cuda.empty_cache()
print(json.dumps({'backbone_bn_layers': sum(s[4] for s in SHAPES), 'total_ms': total_us * 1e-3, 'algorithmic_GB': total_bytes * 1e-9,
                  'mean_GBps': total_bytes / total_us * 1e-3, 'replicas_knob': os.environ.get('PPY_BN_REPLICAS'), 'layers': out}))
