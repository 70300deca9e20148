#!/bin/bash
# quick check: f16x2 kernel + engine parity tests, then the quick bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_f16x2.py tests/test_gpu_engine.py -q -p no:cacheprovider -k "pair or f16x2" > gpurun_out/t_quick.log 2>&1
tail -3 gpurun_out/t_quick.log
timeout 600 python bench.py --steps 20 --warmup 5 --quick --precision f16x2 > gpurun_out/bench_q_f16x2.log 2>&1; tail -1 gpurun_out/bench_q_f16x2.log | cut -c1-260
