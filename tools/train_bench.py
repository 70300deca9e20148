#!/usr/bin/env python
"""BASELINE config C4: ppyolo_2x 608x608, bs=8/GPU, train.py step (frozen-backbone forward with batch-stat BN, head
forward+backward, 6 losses, NCCL gradient all-reduce, fused SGD) on synthetic data.  Prints one JSON line on rank 0.

    python tools/train_bench.py [--steps 10] [--warmup 3] [--precision fp32|bf16] [--size 608] [--batch 8]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench.py ...
"""
import argparse, json, os, sys, time
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, os.path.join(REPO, 'pytorch-ppyolo_b200')); sys.path.insert(0, REPO)
import torch
import torch.distributed as dist
import config as cfgs
from model.ppyolo import PPYOLO
from ppyolo_b200 import synth, targets as tg, parallel
from ppyolo_b200.trainer import Trainer

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=10); ap.add_argument('--warmup', type=int, default=3)
ap.add_argument('--precision', default='fp32'); ap.add_argument('--size', type=int, default=608)
ap.add_argument('--batch', type=int, default=8); ap.add_argument('--arch', default='r50vd')
ap.add_argument('--freeze-at', type=int, default=None, help='override cfg.backbone freeze_at (< 5: trainable backbone stages, autograd_backbone path)')
ap.add_argument('--profile', type=int, default=0, help='also print the top-N kernels of 3 steps (torch.profiler, CUDA time)')
a = ap.parse_args()
rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
cfg = {'r50vd': cfgs.PPYOLO_2x_Config, 'r18vd': cfgs.PPYOLO_r18vd_Config}[a.arch]()
iou_loss = cfgs.select_loss(cfg.iou_loss_type)(**cfg.iou_loss)
iou_aware = cfgs.select_loss(cfg.iou_aware_loss_type)(**cfg.iou_aware_loss) if cfg.head['iou_aware'] else None
yolo = cfgs.select_loss(cfg.yolo_loss_type)(iou_loss=iou_loss, iou_aware_loss=iou_aware, **cfg.yolo_loss)
bb_kw = dict(cfg.backbone)
if a.freeze_at is not None:
    bb_kw['freeze_at'] = a.freeze_at
backbone = cfgs.select_backbone(cfg.backbone_type)(**bb_kw)
head = cfgs.select_head(cfg.head_type)(yolo_loss=yolo, is_train=True, nms_cfg=cfg.nms_cfg, **cfg.head)
model = PPYOLO(backbone, head)
synth.randomize_(model, seed=0)
model.train(); backbone.freeze()
model = model.to(dev)
model.train_precision = a.precision
model.train_graph = bool(int(os.environ.get('PPY_TRAIN_GRAPH', '1')))
trainer = Trainer(model, cfg)
x = synth.images(a.batch, a.size, seed=20 + rank).to(dev)
gb, gc, gs = tg.synthetic_ground_truth(a.batch, seed=30 + rank)
targets = [torch.from_numpy(t).to(dev) for t in tg.gt2yolo_target(gb, gc, gs, h=a.size, w=a.size, **cfg.gt2YoloTarget)]
gb, gc, gs = (torch.from_numpy(v).to(dev) for v in (gb, gc, gs))
for _ in range(a.warmup):
    losses = trainer.step(x, gb, gc, gs, targets)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
t_cpu = time.perf_counter()
for _ in range(a.steps):
    losses = trainer.step(x, gb, gc, gs, targets)
t_cpu = (time.perf_counter() - t_cpu) / a.steps * 1e3          # host time to ENQUEUE a step (no sync inside)
e1.record()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
ms = parallel.max_over_ranks(e0.elapsed_time(e1), dev) / a.steps
sys.stderr.write('rank %d host ms per phase (zero+bind, forward, backward, exchange+tail): %s\n' % (rank, [round(v, 2) for v in trainer.host_ms]))
sys.stderr.write('rank %d timing %s\n' % (rank, json.dumps({k: v for k, v in trainer.timing_summary(8).items() if 'note' not in k})))
if rank == 0:
    print(json.dumps({'metric': 'train_images_per_sec', 'value': world * a.batch / (ms * 1e-3), 'unit': 'images/s', 'n_gpus': world,
                      'ms_per_step': ms, 'cpu_enqueue_ms_per_step': t_cpu, 'steps': a.steps, 'warmup': a.warmup, 'scaling': 'weak',
                      'config': {'workload': 'ppyolo_2x %dx%d bs=%d/GPU train step (freeze_at=%d)' % (a.size, a.size, a.batch, bb_kw['freeze_at']),
                                 'backbone_precision': a.precision, 'trainable_params': int(sum(p.numel() for p in trainer.params)),
                                 'allreduce_bytes': int(trainer.bucket.flat.numel() * 4), 'head_convs': model.train_head_impl or ('kernels (tcgen05 fwd/dgrad/wgrad)' if a.precision == 'bf16' else 'aten (TF32)')},
                      'losses': {k: float(v) for k, v in losses.items()}}), flush=True)
if a.profile and rank == 0:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            trainer.step(x, gb, gc, gs, targets)
        torch.cuda.synchronize()
    rows = {}
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            r = rows.setdefault(ev.name, [0.0, 0])
            r[0] += ev.device_time / 3e3
            r[1] += 1
    total = sum(r[0] for r in rows.values())
    print('profile: %.3f ms of kernels per step, %d distinct kernels' % (total, len(rows)))
    for name, (ms_k, cnt) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:a.profile]:
        print('  %8.3f ms  %5d x  %s' % (ms_k, cnt // 3, name[:150]))
if world > 1:
    dist.destroy_process_group()
