"""Runs eager (non-graph) forward steps of the bench workload so ncu sees every kernel launch by name.
usage: python tools/profile_step.py [--batch 32] [--size 608] [--steps 2] [--arch r50vd]"""
import argparse, os, sys
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, os.path.join(REPO, 'pytorch-ppyolo_b200')); sys.path.insert(0, REPO)
import torch
import bench
from ppyolo_b200 import synth, _lib
from ppyolo_b200.engine import InferenceEngine
ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=32); ap.add_argument('--size', type=int, default=608)
ap.add_argument('--steps', type=int, default=2); ap.add_argument('--arch', default='r50vd')
ap.add_argument('--precision', default='bf16')
a = ap.parse_args()
model, cfg = bench.build_model(a.arch)
model = model.cuda()
eng = InferenceEngine(model, a.batch, a.size, a.size, precision=a.precision, use_graph=False)
eng.x_in.copy_(synth.images(a.batch, a.size, seed=10)); eng.im_size.copy_(synth.im_sizes(a.batch))
torch.cuda.synchronize()
print('launches per step', eng.launches_per_run, 'steps in plan', len(eng.steps))
with open(os.path.join(REPO, 'gpurun_out', 'plan_names.txt'), 'w') as f:
    f.write('\n'.join(n for n, _ in eng.steps))
torch.cuda.cudart().cudaProfilerStart()
for _ in range(a.steps):
    eng.launch()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('done; total launches', _lib.launch_count())
