"""Data-parallel training step (reference train.py:264-285 setup, :427-442 loop body, :460-478 checkpoints).

One process per GPU; every rank runs forward + backward on its own shard of the batch; the ONE exchange step per
iteration is the gradient all-reduce(sum)/world of the trainable (head) parameters -- a single flat fp32 bucket
(92.6 MB for ppyolo_2x) reduced by NCCL over NVLink.  Autograd accumulates every gradient DIRECTLY into its slice of
that bucket (``p.grad`` is a view of it: no pack copies), and the whole optimizer step -- torch.optim.SGD momentum update
with the reference's per-layer lr / weight-decay groups, plus the ExponentialMovingAverage of the new weights
(model/EMA.py:31-45, which the reference round-trips through host memory every step) -- is ONE kernel launch
(``ppy_sgd_ema_multi``) reading the reduced bucket in place.  When the ranks of a node can map each other's buckets (symmetric
memory), the all-reduce itself moves INTO that launch (``ppy_allreduce_sgd_ema``, csrc/allreduce_sgd.cu): barrier on peer signal
pads, in-switch reduction of each rank's slice (``multimem.ld_reduce`` / ``multimem.st`` over NVSwitch, or P2P loads / stores),
barrier, optimizer -- NCCL then only carries the rendezvous.  BatchNorm statistics, DropBlock RNG and data sharding stay
per rank; initial parameters / buffers are broadcast from rank 0; the learning rate follows the
reference's warm-up + piecewise decay and is NOT rescaled by the world size, like the reference."""
import ctypes
import os

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
    """Flat fp32 buffer holding every gradient: per-parameter views, all_reduce(sum) over the ranks (device agnostic)."""

    def __init__(self, params):
        self.params = list(params)
        self.sizes = [p.numel() for p in self.params]
        total = sum(self.sizes)
        dev = self.params[0].device if self.params else 'cpu'
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.offsets = [0]
        for s in self.sizes:
            self.offsets.append(self.offsets[-1] + s)

    def rebase(self, storage):
        """Move the bucket into caller-provided storage (symmetric memory mapped into every peer): same length, same views."""
        assert storage.numel() >= self.flat.numel() and storage.dtype == torch.float32
        storage[:self.flat.numel()].copy_(self.flat)
        self.flat = storage[:self.flat.numel()]

    def view(self, i):
        return self.flat[self.offsets[i]:self.offsets[i + 1]]

    def bind_grads(self):
        """Make every ``p.grad`` a view of the bucket, so backward accumulates straight into it."""
        for i, p in enumerate(self.params):
            p.grad = self.view(i).view_as(p)

    def pack(self):
        """Copy gradients that live elsewhere (a caller replaced ``p.grad``) into the bucket; bound views cost nothing."""
        for i, p in enumerate(self.params):
            v = self.view(i)
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad.reshape(-1))

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
    def __init__(self, model, cfg, graph=None, ema=None):
        """``graph``: replay the frozen-backbone forward and the head forward + losses + backward as CUDA graphs (one set per
        input shape); None = keep the model's ``train_graph`` setting.  ``ema``: keep the reference's exponential moving average of
        the trainable weights (None = ``cfg.use_ema``), updated inside the optimizer kernel."""
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
        rank, world = parallel.world()
        if world > 1:                      # replicas must start from the same weights and statistics
            with torch.no_grad():
                for t in list(model.parameters()) + list(model.buffers()):
                    dist.broadcast(t.data, 0)
        self.bucket = GradientBucket(self.params)
        dev = self.bucket.flat.device
        self.momentum_flat = torch.zeros_like(self.bucket.flat)
        self.momentum_bufs = [self.momentum_flat[self.bucket.offsets[i]:self.bucket.offsets[i + 1]].view_as(p)
                              for i, p in enumerate(self.params)]
        self.iter_id = 0
        self.ema = None
        use_ema = getattr(cfg, 'use_ema', False) if ema is None else ema
        self._cuda = dev.type == 'cuda'
        if self._cuda:
            self._offsets_dev = torch.tensor(self.bucket.offsets, dtype=torch.int64, device=dev)
            self._lr_mult = torch.tensor([g['base_lr'] / self.base_lr for g in groups], dtype=torch.float32, device=dev)
            self._wd = torch.tensor([g['weight_decay'] for g in groups], dtype=torch.float32, device=dev)
            self._ptr_key, self._ptr_table = None, None
            if use_ema:
                from model.EMA import ExponentialMovingAverage
                self.ema = ExponentialMovingAverage(model, cfg.ema_decay)
                self.ema.register()
                at = {id(p): self.ema._offsets[i] for i, p in enumerate(self.ema._params)}
                self._shadow_offsets = torch.tensor([at[id(p)] for p in self.params], dtype=torch.int64, device=dev)
        self._events = []
        # the exchange step as ONE kernel over NVLink peer memory (csrc/allreduce_sgd.cu) when the ranks can map each other's
        # gradient buckets; NCCL all-reduce + the optimizer kernel otherwise (PPY_PEER_ALLREDUCE=0 forces that)
        self._peer, self.exchange_impl = None, 'nccl all-reduce + optimizer kernel' if world > 1 else 'single rank: optimizer kernel'
        if world > 1 and self._cuda and os.environ.get('PPY_PEER_ALLREDUCE', '1') != '0':
            self._setup_peer_exchange(rank, world)

    def _setup_peer_exchange(self, rank, world):
        """Put the gradient bucket into symmetric memory (torch.distributed._symmetric_memory: CUDA VMM allocations mapped into every
        rank of the node, NVLS multicast address when the switch offers one) and keep the peer pointer tables for the kernel.  All
        ranks switch together or not at all."""
        dev = self.bucket.flat.device
        state, err = None, ''
        try:
            import torch.distributed._symmetric_memory as symm_mem
            total = self.bucket.flat.numel()
            padded = (total + 4 * world - 1) // (4 * world) * (4 * world)
            store = symm_mem.empty(padded, dtype=torch.float32, device=dev)
            store.zero_()
            hdl = symm_mem.rendezvous(store, dist.group.WORLD)
            if hdl.world_size != world or hdl.rank != rank or hdl.signal_pad_size < 4 * (self.PAD_SLOT + world):
                raise RuntimeError('symmetric memory handle does not match the process group')
            mc = 0 if os.environ.get('PPY_PEER_NO_MULTICAST') else int(hdl.multicast_ptr or 0)      # knob: exercise the P2P form on an NVLS box
            state = {'store': store, 'hdl': hdl, 'padded': padded, 'multicast': mc,
                     'peers': int(hdl.buffer_ptrs_dev), 'pads': int(hdl.signal_pad_ptrs_dev),
                     'done': torch.zeros(2, dtype=torch.int32, device=dev)}
        except Exception as exc:                # no symmetric memory on this box / torch build: NCCL path
            err = '%s: %s' % (type(exc).__name__, str(exc)[:120])
        ok = torch.tensor([1 if state is not None else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 1:
            self.bucket.rebase(state['store'])
            self._peer = state
            self.exchange_impl = 'one kernel per rank over peer memory: %s all-reduce + SGD + EMA (csrc/allreduce_sgd.cu)' % (
                'NVLS multimem.ld_reduce / multimem.st' if state['multicast'] else 'P2P load / store')
        elif err:
            self.exchange_impl += ' (peer-memory kernel unavailable: %s)' % err

    PAD_SLOT = 2048          # first uint32 word of the signal pads this trainer uses (torch's own barriers use the low channels)

    # ------------------------------------------------------------------ one iteration
    def _pointer_table(self):
        ptrs = [p.data_ptr() for p in self.params]
        if self._ptr_key != ptrs:                                  # EMA.apply()/restore() rebind param.data
            self._ptr_key, self._ptr_table = ptrs, torch.tensor(ptrs, dtype=torch.int64, device=self.bucket.flat.device)
        return self._ptr_table

    def step(self, images, gt_bbox, gt_class, gt_score, targets):
        from ._lib import lib, check
        from . import ops
        if not self._cuda:
            raise RuntimeError('ppyolo_b200: the training step needs CUDA parameters (no CPU fallback)')
        import time
        tc = [time.perf_counter()]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        self.bucket.flat.zero_()
        self.bucket.bind_grads()
        tc.append(time.perf_counter())
        losses = self.model(images, None, False, gt_bbox, gt_class, gt_score, targets)
        tc.append(time.perf_counter())
        total = sum(losses.values())
        total.backward()
        tc.append(time.perf_counter())
        self.bucket.pack()                                         # no-op for the bound views
        ev[1].record()
        grad_scale = 1.0 / parallel.world()[1] if self._peer is not None else self.bucket.all_reduce()
        ev[2].record()
        lr = calc_lr(self.iter_id, self.cfg)
        first = 1 if self.iter_id == 0 else 0
        shadow, sh_off, decay = None, None, 0.0
        if self.ema is not None:
            import numpy as np
            step = self.ema._update_step
            decay = min(self.ema._decay, (1 + step) / (10 + step)) if self.ema._thres_steps else self.ema._decay
            shadow, sh_off = self.ema._shadow_flat, self._shadow_offsets
            d32, omd32 = float(np.float32(decay)), float(np.float32(1 - decay))
        else:
            d32, omd32 = 0.0, 0.0
        if self._peer is not None:
            pe = self._peer
            rank, world = parallel.world()
            check(lib.ppy_allreduce_sgd_ema(ops.ptr(self.bucket.flat), ctypes.c_void_p(pe['multicast'] or None), ctypes.c_void_p(pe['peers']),
                                            ctypes.c_void_p(pe['pads']), self.PAD_SLOT, rank, world, self.iter_id + 1, ops.ptr(pe['done']),
                                            pe['padded'], ops.ptr(self._pointer_table()), ops.ptr(self.momentum_flat), ops.ptr(shadow),
                                            ops.ptr(self._offsets_dev), ops.ptr(sh_off), ops.ptr(self._lr_mult), ops.ptr(self._wd),
                                            len(self.params), float(lr), float(self.momentum), float(grad_scale), first, d32, omd32,
                                            ops.stream_ptr()), 'allreduce_sgd_ema')
        else:
            check(lib.ppy_sgd_ema_multi(ops.ptr(self._pointer_table()), ops.ptr(self.bucket.flat), ops.ptr(self.momentum_flat), ops.ptr(shadow),
                                        ops.ptr(self._offsets_dev), ops.ptr(sh_off), ops.ptr(self._lr_mult), ops.ptr(self._wd), len(self.params),
                                        float(lr), float(self.momentum), float(grad_scale), first, d32, omd32, ops.stream_ptr()), 'sgd_ema_multi')
        ev[3].record()
        if self.ema is not None:
            self.ema._update_step += 1
        with torch.no_grad():
            # the kernel wrote through raw pointers: move torch's version counters too.  ONE call with the list -- the binding
            # takes an iterable of tensors, and handing it a single tensor makes it iterate over (unbind) that tensor's rows:
            # 16 ms of host time per step, the whole difference between a host-bound and a GPU-bound iteration
            torch._C._increment_version(self.params)
        self.iter_id += 1
        if self._peer is not None and not self._peer.get('checked'):
            # first step through the peer-memory kernel: make sure every rank arrived at its barriers before trusting the path
            torch.cuda.synchronize()
            self._peer['checked'] = True
            self.check_exchange()
        self.model.invalidate_engines_for_weights()
        tc.append(time.perf_counter())
        self.host_ms = [(b - a) * 1e3 for a, b in zip(tc[:-1], tc[1:])]     # host time: zero+bind / forward / backward / exchange+tail
        self._events.append(ev)
        if len(self._events) > 64:
            self._events.pop(0)
        return {k: v.detach() for k, v in losses.items()}

    def timing_summary(self, last=10):
        """Mean CUDA-event times of the last steps: forward+backward / gradient all-reduce / optimizer(+EMA) kernel."""
        torch.cuda.synchronize()
        self.check_exchange()
        evs = self._events[-last:]
        if not evs:
            return {}
        mean = lambda a, b: sum(e[a].elapsed_time(e[b]) for e in evs) / len(evs)
        ar = mean(1, 2)
        out = {'fwd_bwd_ms': mean(0, 1), 'allreduce_ms': ar, 'optimizer_ema_ms': mean(2, 3), 'allreduce_overlap_fraction': 0.0,
               'exchange_impl': self.exchange_impl, 'ema': self.ema is not None}
        if self._peer is not None:
            out['allreduce_ms'] = None
            out['exchange_optimizer_ms'] = out.pop('optimizer_ema_ms')
            out['allreduce_note'] = ('the all-reduce runs INSIDE the optimizer launch (barrier on peer signal pads, in-switch reduction of '
                                     "the rank's slice, barrier, optimizer): exchange_optimizer_ms is all-reduce + SGD + EMA")
        else:
            out['allreduce_note'] = 'one NCCL all-reduce of the flat bucket after the (single CUDA graph) backward; not overlapped'
        return out

    def check_exchange(self):
        """Raise if a barrier of the peer-memory exchange kernel ever timed out (a rank did not make the matching call)."""
        if self._peer is not None and int(self._peer['done'][1].item()) != 0:
            raise RuntimeError('ppyolo_b200: the peer-memory gradient exchange timed out waiting for a rank; parameters are not in sync')

    # ------------------------------------------------------------------ checkpoint / resume
    def state_dict(self):
        """Optimizer + EMA state next to the model weights (the reference saves the weights only, train.py:460-478, so a resumed
        run restarts momentum and EMA from scratch)."""
        self.check_exchange()
        sd = {'iter_id': self.iter_id, 'momentum_flat': self.momentum_flat.clone(),
              'param_shapes': [tuple(p.shape) for p in self.params]}
        if self.ema is not None:
            sd['ema_shadow_flat'] = self.ema._shadow_flat.clone()
            sd['ema_update_step'] = self.ema._update_step
            sd['ema_names'] = list(self.ema._names)
        return sd

    def load_state_dict(self, sd):
        if [tuple(s) for s in sd['param_shapes']] != [tuple(p.shape) for p in self.params]:
            raise ValueError('trainer checkpoint does not match the trainable parameters of this model')
        self.iter_id = int(sd['iter_id'])
        self.momentum_flat.copy_(sd['momentum_flat'])
        if self.ema is not None and 'ema_shadow_flat' in sd:
            if list(sd['ema_names']) != list(self.ema._names):
                raise ValueError('EMA checkpoint tensor names differ')
            self.ema._shadow_flat.copy_(sd['ema_shadow_flat'])
            self.ema._update_step = int(sd['ema_update_step'])

    def save_checkpoint(self, directory, keep=10):
        """``step%08d.pt`` = model.state_dict() exactly like the reference (train.py:462-463; resume = iteration parsed from the
        file name, :255-261) + ``step%08d.opt.pt`` with the optimizer / EMA state; keeps the newest ``keep`` (train.py:464-478)."""
        os.makedirs(directory, exist_ok=True)
        path = os.path.join(directory, 'step%.8d.pt' % self.iter_id)
        torch.save(self.model.state_dict(), path)
        torch.save(self.state_dict(), path[:-3] + '.opt.pt')
        names = sorted(f for f in os.listdir(directory) if f.startswith('step') and f.endswith('.pt') and not f.endswith('.opt.pt'))
        for old in names[:-keep] if keep > 0 else []:
            for f in (old, old[:-3] + '.opt.pt'):
                try:
                    os.remove(os.path.join(directory, f))
                except OSError:
                    pass
        return path

    def load_checkpoint(self, path):
        self.model.load_state_dict(torch.load(path, map_location=self.bucket.flat.device))
        opt = path[:-3] + '.opt.pt'
        if os.path.exists(opt):
            self.load_state_dict(torch.load(opt, map_location=self.bucket.flat.device))
        else:                                   # a reference-style checkpoint: iteration from the file name (train.py:258-260)
            base = os.path.basename(path)
            if base.startswith('step') and base[4:12].isdigit():
                self.iter_id = int(base[4:12])
        return self.iter_id
