#!/bin/bash
# ncu launch list of the library's kernels in one eager training step (config C4)
mkdir -p gpurun_out
PPY_TRAIN_GRAPH=0 timeout 800 ncu --metrics gpu__time_duration.sum,launch__grid_size,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --cache-control none \
   -k regex:"conv_umma|kmajor|dcn_umma|bn_train|bn_act|yolo_loss|pack_weight|sgd_ema|dropblock|spp_|pool|stem|gt2yolo|nchw" --csv --log-file gpurun_out/train_launches_all.csv \
   python tools/train_bench.py --precision bf16 --steps 1 --warmup 2 > gpurun_out/train_ncu3.log 2>&1
tail -1 gpurun_out/train_ncu3.log | cut -c1-120; wc -l gpurun_out/train_launches_all.csv
