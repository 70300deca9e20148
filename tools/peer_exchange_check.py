#!/usr/bin/env python
"""Two or more ranks (torchrun): the training step's exchange as ONE peer-memory kernel (ppy_allreduce_sgd_ema: in-switch all-reduce +
SGD + EMA) against the NCCL all-reduce + optimizer kernel path, from identical initial states on identical per-rank data:
parameters, momentum and EMA shadows after 3 steps must agree (summation order differs: 1e-6 relative) and be bit-identical across
the ranks.  Prints one JSON line on rank 0."""
import copy, json, os, sys
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, os.path.join(REPO, 'pytorch-ppyolo_b200')); sys.path.insert(0, REPO)
import torch
import torch.distributed as dist
from tests.test_gpu_train import build_train_model, train_inputs
from ppyolo_b200.trainer import Trainer

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
import tests.test_gpu_train as tt
tt.DEV = dev


def run(peer):
    os.environ['PPY_PEER_ALLREDUCE'] = '1' if peer else '0'
    model, cfg = build_train_model('r50vd')
    model.train_precision = 'bf16'
    trainer = Trainer(model, cfg, graph=False, ema=True)
    x, gb, gc, gs, targets = train_inputs(cfg, size=128, batch=2)
    x = x + 0.01 * rank                                    # different data per rank: the reduction matters
    for _ in range(3):
        losses = trainer.step(x, gb, gc, gs, targets)
    torch.cuda.synchronize()
    flat = torch.cat([p.detach().flatten() for p in trainer.params])
    return trainer, flat, trainer.momentum_flat.clone(), trainer.ema._shadow_flat.clone(), {k: float(v) for k, v in losses.items()}


t_peer, p_peer, m_peer, s_peer, l_peer = run(True)
t_nccl, p_nccl, m_nccl, s_nccl, l_nccl = run(False)


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def same_on_all_ranks(t):
    got = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(got, t.contiguous())
    return all(bool((g == got[0]).all()) for g in got)


out = {'world': world, 'peer_impl': t_peer.exchange_impl, 'nccl_impl': t_nccl.exchange_impl,
       'param_rel_err': rel(p_peer, p_nccl), 'momentum_rel_err': rel(m_peer, m_nccl), 'shadow_rel_err': rel(s_peer, s_nccl),
       'params_bit_identical_across_ranks': same_on_all_ranks(p_peer), 'momentum_bit_identical_across_ranks': same_on_all_ranks(m_peer),
       'losses_peer': l_peer, 'losses_nccl': l_nccl, 'timing_peer': t_peer.timing_summary(2), 'timing_nccl': t_nccl.timing_summary(2)}
ok = t_peer._peer is not None and out['param_rel_err'] < 1e-5 and out['momentum_rel_err'] < 1e-4 and out['params_bit_identical_across_ranks']
out['ok'] = bool(ok)
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
