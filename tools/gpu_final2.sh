#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err
tail -1 gpurun_out/bench_full.log | cut -c1-200
