#!/bin/bash
# DCN v2 kernel: parity tests, micro-benchmark, quick bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_f16x2.py -q -s -p no:cacheprovider -k "whole_layer or pair_dcn" > gpurun_out/t_dcn2.log 2>&1
tail -4 gpurun_out/t_dcn2.log
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_engine.py -q -s -p no:cacheprovider -k "dcn or golden or detections" > gpurun_out/t_dcn3.log 2>&1
tail -5 gpurun_out/t_dcn3.log; grep "far offsets\|detections" gpurun_out/t_dcn3.log | cut -c1-250
timeout 300 python tools/dcn_bench.py > gpurun_out/dcn_bench_new.json 2>&1; tail -1 gpurun_out/dcn_bench_new.json
timeout 600 python bench.py --steps 10 --warmup 3 --quick --precision f16x2 > gpurun_out/bench_q_f16x2.log 2>&1; tail -1 gpurun_out/bench_q_f16x2.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --quick --precision bf16 > gpurun_out/bench_q_bf16.log 2>&1; tail -1 gpurun_out/bench_q_bf16.log | cut -c1-300
