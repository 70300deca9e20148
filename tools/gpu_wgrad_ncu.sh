#!/bin/bash
mkdir -p gpurun_out
PPY_TRAIN_GRAPH=0 timeout 500 ncu --set full --import-source on --clock-control none --cache-control none \
   -k regex:"conv_umma|kmajor|dcn_umma" --launch-skip 734 --launch-count 1 -o gpurun_out/prof_wgrad -f \
   python tools/train_bench.py --precision bf16 --steps 1 --warmup 3 > gpurun_out/train_ncu2.log 2>&1
tail -2 gpurun_out/train_ncu2.log | cut -c1-200
