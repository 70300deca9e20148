#!/bin/bash
# training-path check: GPU tests of the training step (losses, targets, optimizer) + the C4 step benchmark
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py tests/test_gpu_dropblock.py -q -p no:cacheprovider > gpurun_out/t_train.log 2>&1
tail -5 gpurun_out/t_train.log
timeout 600 python tools/train_bench.py --precision bf16 --steps 10 --warmup 3 --profile 45 > gpurun_out/train_bench.log 2>&1; grep -v Warning gpurun_out/train_bench.log | tail -50 | cut -c1-400
