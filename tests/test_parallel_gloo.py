"""world_size-2 gloo test of the batch-sharding host logic (the N>1 path has no data-path collective)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ppyolo_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, ws, port, total, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    try:
        sl = parallel.image_shard(total)
        # stand-in per-image detections: row count and content identify the global image index
        local = [torch.full((i % 3 + 1, 6), float(i)) for i in range(total)[sl]]
        merged = parallel.gather_detections(local)
        slow = parallel.max_over_ranks(10.0 + rank)
        dist.barrier()
        torch.save({'slice': (sl.start, sl.stop), 'merged': merged, 'slow': slow}, os.path.join(out_dir, 'r%d.pt' % rank))
    finally:
        dist.destroy_process_group()


def test_shard_partition_is_exact():
    for total in (1, 7, 32, 33):
        for ws in (1, 2, 4, 8):
            seen = []
            for r in range(ws):
                sl = parallel.image_shard(total, r, ws)
                seen += list(range(total))[sl]
            assert seen == list(range(total))


def test_two_rank_gloo(tmp_path):
    total, ws = 7, 2
    mp.spawn(_worker, args=(ws, _free_port(), total, str(tmp_path)), nprocs=ws, join=True)
    res = [torch.load(os.path.join(str(tmp_path), 'r%d.pt' % r)) for r in range(ws)]
    assert res[0]['slice'] == (0, 4) and res[1]['slice'] == (4, 7)
    for r in res:
        assert r['slow'] == 11.0
        assert [int(p[0, 0]) for p in r['merged']] == list(range(total))
        assert [p.shape[0] for p in r['merged']] == [i % 3 + 1 for i in range(total)]


def _grad_worker(rank, ws, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    try:
        from ppyolo_b200.trainer import GradientBucket
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2, 2, 2))]
        for i, p in enumerate(params):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        params[1].grad = None if rank == 1 else params[1].grad      # a rank without a gradient contributes zeros
        bucket = GradientBucket(params)
        bucket.pack()
        scale = bucket.all_reduce()
        bucket.unpack_mean(scale)
        torch.save({'grads': [p.grad.clone() for p in params], 'scale': scale}, os.path.join(out_dir, 'g%d.pt' % rank))
    finally:
        dist.destroy_process_group()


def test_gradient_bucket_allreduce_gloo(tmp_path):
    """The one exchange step of training: flat-bucket all-reduce(sum)/world, world_size 2 on gloo."""
    ws = 2
    mp.spawn(_grad_worker, args=(ws, _free_port(), str(tmp_path)), nprocs=ws, join=True)
    res = [torch.load(os.path.join(str(tmp_path), 'g%d.pt' % r)) for r in range(ws)]
    for r in res:
        assert r['scale'] == 0.5
        assert torch.allclose(r['grads'][0], torch.full((3, 4), 1.5))       # (1 + 2) / 2
        assert torch.allclose(r['grads'][1], torch.full((5,), 1.0))         # (2 + 0) / 2
        assert torch.allclose(r['grads'][2], torch.full((2, 2, 2), 4.5))    # (3 + 6) / 2
