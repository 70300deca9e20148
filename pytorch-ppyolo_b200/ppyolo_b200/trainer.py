"""Data-parallel training step (reference train.py:264-285 setup, :427-442 loop body).

One process per GPU; every rank runs forward + backward on its own shard of the batch; the ONE exchange step per
iteration is the gradient all-reduce(sum)/world of the trainable (head) parameters -- a single flat fp32 bucket
(92.6 MB for ppyolo_2x) reduced by NCCL over NVLink -- followed by a fused SGD-momentum kernel that reads the reduced
bucket in place.  BatchNorm statistics, DropBlock RNG and data sharding stay per rank; the learning
rate follows the reference's warm-up + piecewise decay and is NOT rescaled by the world size, like the reference."""
import ctypes

import torch
import torch.distributed as dist

from . import parallel


def calc_lr(iter_id, cfg):
    """Reference train.py:172-188."""
    base_lr = cfg.learningRate['base_lr']
    decay, warm = cfg.learningRate['PiecewiseDecay'], cfg.learningRate['LinearWarmup']
    milestones = decay['milestones']
    for i in range(len(milestones), 0, -1):
        if iter_id >= milestones[i - 1]:
            return base_lr * decay['gamma'] ** i
    if iter_id <= warm['steps']:
        return base_lr * (warm['start_factor'] + (1.0 - warm['start_factor']) / warm['steps'] * iter_id)
    return base_lr


class GradientBucket(object):
    """Flat fp32 view of a list of gradients: pack -> all_reduce(sum) -> per-parameter views (device agnostic)."""

    def __init__(self, params):
        self.params = list(params)
        self.sizes = [p.numel() for p in self.params]
        total = sum(self.sizes)
        dev = self.params[0].device if self.params else 'cpu'
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.offsets = [0]
        for s in self.sizes:
            self.offsets.append(self.offsets[-1] + s)

    def view(self, i):
        return self.flat[self.offsets[i]:self.offsets[i + 1]]

    def pack(self):
        for i, p in enumerate(self.params):
            if p.grad is None:
                self.view(i).zero_()
            else:
                self.view(i).copy_(p.grad.reshape(-1))

    def all_reduce(self):
        """Sum over ranks; returns the factor the consumer must apply (1/world)."""
        rank, world = parallel.world()
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        return 1.0 / world

    def unpack_mean(self, scale):
        for i, p in enumerate(self.params):
            p.grad = (self.view(i) * scale).reshape(p.shape).clone()


class Trainer(object):
    def __init__(self, model, cfg, graph=None):
        """``graph``: replay the frozen-backbone forward and the head forward + losses + backward as CUDA graphs (one set per
        input shape); None = keep the model's ``train_graph`` setting."""
        self.model, self.cfg = model, cfg
        if graph is not None:
            model.train_graph = bool(graph)
        self.base_lr = cfg.learningRate['base_lr']
        self.base_wd = cfg.optimizerBuilder['regularizer']['factor']
        self.momentum = cfg.optimizerBuilder['optimizer']['momentum']
        groups = []
        model.add_param_group(groups, self.base_lr, self.base_wd)        # per-tensor groups, reference custom_layers.py:167-241
        self.groups = groups
        self.params = [g['params'][0] for g in groups]
        self.bucket = GradientBucket(self.params)
        self.momentum_bufs = [torch.zeros_like(p, dtype=torch.float32) for p in self.params]
        self.iter_id = 0

    def step(self, images, gt_bbox, gt_class, gt_score, targets):
        from ._lib import lib, check
        from . import ops
        losses = self.model(images, None, False, gt_bbox, gt_class, gt_score, targets)
        total = sum(losses.values())
        for p in self.params:
            p.grad = None
        total.backward()
        self.bucket.pack()
        grad_scale = self.bucket.all_reduce()
        lr = calc_lr(self.iter_id, self.cfg)
        first = 1 if self.iter_id == 0 else 0
        for i, (g, p) in enumerate(zip(self.groups, self.params)):
            if not p.is_cuda:
                raise RuntimeError('ppyolo_b200: the fused SGD kernel needs CUDA parameters')
            group_lr = lr * g['base_lr'] / self.base_lr
            gv = self.bucket.view(i)
            check(lib.ppy_sgd_momentum(ops.ptr(p.data), ctypes.c_void_p(gv.data_ptr()), ops.ptr(self.momentum_bufs[i]), p.numel(),
                                       float(group_lr), float(self.momentum), float(g['weight_decay']), float(grad_scale), first,
                                       ops.stream_ptr()), 'sgd_momentum')
        for p in self.params:                 # the kernel wrote through raw pointers: move torch's version counters too
            torch._C._increment_version(p)
        self.iter_id += 1
        self.model.invalidate_engines_for_weights()
        return {k: v.detach() for k, v in losses.items()}
