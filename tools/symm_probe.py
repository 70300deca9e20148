"""Probe: does torch symmetric memory (CUDA VMM peer mappings, NVLS multicast) work on this box?  torchrun --nproc-per-node N."""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
t.fill_(rank + 1)
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
info = {}
for name in ('world_size', 'rank', 'multicast_ptr', 'buffer_ptrs', 'signal_pad_ptrs', 'signal_pad_size', 'buffer_size', 'buffer_ptrs_dev', 'signal_pad_ptrs_dev'):
    try:
        v = getattr(hdl, name)
        info[name] = [hex(p) for p in v] if isinstance(v, (list, tuple)) else (hex(v) if isinstance(v, int) and v > 4096 else v)
    except Exception as e:
        info[name] = 'ERR ' + repr(e)[:80]
print(rank, info, flush=True)
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (4,), torch.float32)
print(rank, 'peer values', peer.tolist(), flush=True)
hdl.barrier()
dist.destroy_process_group()
