#!/bin/bash
# whole GPU suite + training-step kernel profile
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1
tail -4 gpurun_out/t_all.log | cut -c1-300
timeout 600 python tools/train_bench.py --precision bf16 --steps 10 --warmup 3 --profile 70 > gpurun_out/train_bench.log 2>&1; grep -v Warning gpurun_out/train_bench.log | tail -75 | cut -c1-330
