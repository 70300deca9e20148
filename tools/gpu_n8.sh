#!/bin/bash
# the driver's N-GPU launch of bench.py (training sub-record: peer-memory exchange kernel at N ranks)
N=${1:-8}
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
echo rc=$?
tail -1 gpurun_out/bench_n$N.log | cut -c1-400
tail -3 gpurun_out/bench_n$N.err | cut -c1-300
