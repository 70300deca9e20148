#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_dropblock.py -q -p no:cacheprovider -s > gpurun_out/t_pre.log 2>&1
grep -E "resize|drift|passed|failed|Error|error" gpurun_out/t_pre.log | tail -30
timeout 600 python tools/train_bench.py --precision bf16 --steps 10 --warmup 3 --profile 12 > gpurun_out/train_bench.log 2>&1; grep -v Warning gpurun_out/train_bench.log | tail -14 | cut -c1-300
