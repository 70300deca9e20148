#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --cache-control none --import-source on -k regex:yolo_loss --launch-count 6 \
   -o gpurun_out/prof_loss -f python -m pytest tests/test_gpu_train.py -q -p no:cacheprovider -x -k test_train_c4_full_shape > gpurun_out/ncu_loss.log 2>&1
tail -3 gpurun_out/ncu_loss.log
