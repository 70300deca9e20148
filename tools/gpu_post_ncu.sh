#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:bn_train_fused --launch-skip 6 --launch-count 1 \
   -o gpurun_out/prof_bn_post -f python tools/bn_bench.py --only stage3.conv1/2 > gpurun_out/ncu_bn_post.log 2>&1
tail -1 gpurun_out/ncu_bn_post.log
PPY_TRAIN_GRAPH=0 timeout 500 ncu --set full --import-source on --clock-control none --cache-control none \
   -k regex:"conv_umma_kernel" --launch-skip 520 --launch-count 40 -o gpurun_out/prof_wgrad_post -f \
   python tools/train_bench.py --precision bf16 --steps 1 --warmup 3 > gpurun_out/train_ncu4.log 2>&1
tail -1 gpurun_out/train_ncu4.log | cut -c1-100
