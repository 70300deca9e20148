#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/t_all.log 2>&1
tail -3 gpurun_out/t_all.log | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err
tail -1 gpurun_out/bench_full.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
