#!/bin/bash
# full GPU test suite + the complete bench line + the reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1
tail -6 gpurun_out/t_all.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.log 2>&1
tail -1 gpurun_out/bench_full.log | cut -c1-1500
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
tail -1 gpurun_out/bench_ref.log | cut -c1-400
