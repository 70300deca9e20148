#!/usr/bin/env python
"""BASELINE configs[4] in isolation: Matrix-NMS on 10k boxes x 80 classes per image through the C ABI, us per image at several
batch sizes (CUDA events; plain launches and CUDA-graph replays).  `--once` runs one call per batch size (for an ncu launch list)."""
import argparse, json, os, sys
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, os.path.join(REPO, 'pytorch-ppyolo_b200')); sys.path.insert(0, REPO)
import torch
from ppyolo_b200 import ops, synth

ap = argparse.ArgumentParser()
ap.add_argument('--once', action='store_true'); ap.add_argument('--batches', default='1,8,32')
a = ap.parse_args()
dev = torch.device('cuda')
b, s = synth.nms_inputs(10000, 80, seed=0)
print('candidates > 0.01 per image:', int((s > 0.01).sum()))
res = {}
for bs in [int(v) for v in a.batches.split(',')]:
    boxes = b[None].repeat(bs, 1, 1).to(dev).contiguous()
    scores = s[None].repeat(bs, 1, 1).to(dev).contiguous()
    out = torch.empty((bs, 100, 6), dtype=torch.float32, device=dev)
    counts = torch.empty((bs,), dtype=torch.int32, device=dev)
    ws = ops.nms_workspace(bs, 10000, 80, dev)
    run = lambda: ops.matrix_nms_launch(boxes, scores, out, counts, ws, 0.01, 0.01, 500, 100, False, 2.0)
    if a.once:
        torch.cuda.synchronize(); torch.cuda.profiler.start(); run(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
        continue
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        run()
    e1.record(); torch.cuda.synchronize()
    plain = e0.elapsed_time(e1) / 50 * 1e3
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        run(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=st):
            run()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    res[bs] = {'us_per_call': plain, 'us_per_img': plain / bs, 'graph_us_per_call': e0.elapsed_time(e1) / 50 * 1e3, 'detections': counts.cpu().tolist()[:2]}
print(json.dumps(res))
