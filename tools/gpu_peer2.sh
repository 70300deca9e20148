#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/peer_exchange_check.py > gpurun_out/peer_check_n$N.log 2>&1
tail -1 gpurun_out/peer_check_n$N.log | cut -c1-1500
