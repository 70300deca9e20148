#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropblock.py tests/test_gpu_train.py -q -p no:cacheprovider -x > gpurun_out/t_train.log 2>&1
tail -4 gpurun_out/t_train.log | cut -c1-300
timeout 600 python tools/train_bench.py --precision bf16 --steps 10 --warmup 3 --profile 40 > gpurun_out/train_bench.log 2>&1; grep -v Warning gpurun_out/train_bench.log | tail -45 | cut -c1-250
