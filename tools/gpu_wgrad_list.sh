#!/bin/bash
mkdir -p gpurun_out
PPY_TRAIN_GRAPH=0 timeout 500 ncu --metrics gpu__time_duration.sum,launch__grid_size,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --cache-control none \
   -k regex:"conv_umma|kmajor|dcn_umma" --launch-skip 700 --launch-count 260 --csv --log-file gpurun_out/train_launches.csv \
   python tools/train_bench.py --precision bf16 --steps 1 --warmup 3 > gpurun_out/train_ncu.log 2>&1
tail -2 gpurun_out/train_ncu.log | cut -c1-200; wc -l gpurun_out/train_launches.csv
