"""Device-resident throughput of the inference engine for any (arch, size, batch): CUDA-graph replays timed with CUDA events.
usage: python tools/engine_bench.py [--arch r18vd] [--size 416] [--batch 16] [--steps 50]   (BASELINE configs[1] by default)"""
import argparse, os, sys
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, os.path.join(REPO, 'pytorch-ppyolo_b200')); sys.path.insert(0, REPO)
import torch
import bench
from ppyolo_b200 import synth
ap = argparse.ArgumentParser()
ap.add_argument('--arch', default='r18vd'); ap.add_argument('--size', type=int, default=416)
ap.add_argument('--batch', type=int, default=16); ap.add_argument('--steps', type=int, default=50)
ap.add_argument('--precision', default='bf16')
a = ap.parse_args()
model, cfg = bench.build_model(a.arch)
model = model.cuda(); model.precision = a.precision
eng = model.engine(a.batch, a.size, a.size)
eng.x_in.copy_(synth.images(a.batch, a.size, seed=10)); eng.im_size.copy_(synth.im_sizes(a.batch))
for _ in range(5): eng.launch()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps): eng.launch()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
flops = eng.conv_flops
print('%s %dx%d bs=%d %s: %.3f ms/step, %.0f img/s, conv %.1f TFLOP/s (%d launches/step)' % (
    a.arch, a.size, a.size, a.batch, a.precision, ms, a.batch / ms * 1e3, flops / ms / 1e9, eng.launches_per_run))
