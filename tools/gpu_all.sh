#!/bin/bash
# whole GPU suite + quick headline bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1
tail -8 gpurun_out/t_all.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --quick --precision f16x2 > gpurun_out/bench_q_f16x2.log 2>&1; tail -1 gpurun_out/bench_q_f16x2.log | cut -c1-260
