#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1
tail -2 gpurun_out/t_all.log | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err
tail -1 gpurun_out/bench_full.log | cut -c1-200
