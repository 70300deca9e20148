#!/bin/bash
# full GPU test suite + full bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/t_all.log 2>&1
tail -15 gpurun_out/t_all.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.log 2>&1
tail -1 gpurun_out/bench_full.log | cut -c1-700
