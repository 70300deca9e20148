#!/bin/bash
# first GPU round trip of round 2: the f16x2 kernel + engine tests, then a bench line per precision
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_f16x2.py -q -s -p no:cacheprovider > gpurun_out/t_f16x2.log 2>&1
tail -25 gpurun_out/t_f16x2.log
timeout 1200 python -m pytest tests/test_gpu_engine.py -q -s -p no:cacheprovider -k "f16x2 or drift" > gpurun_out/t_engine.log 2>&1
tail -25 gpurun_out/t_engine.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision f16x2 > gpurun_out/bench_f16x2.log 2>&1
cp gpurun_out/per_op_ms.json gpurun_out/per_op_ms_f16x2.json
tail -1 gpurun_out/bench_f16x2.log | cut -c1-600
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision bf16 > gpurun_out/bench_bf16.log 2>&1
tail -1 gpurun_out/bench_bf16.log | cut -c1-400
