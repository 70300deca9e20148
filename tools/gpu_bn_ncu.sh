#!/bin/bash
mkdir -p gpurun_out
for L in stage3.conv1/2 stage4.conv1/2; do
  T=$(echo $L | tr '/.' '__')
  timeout 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:bn_train_fused --launch-skip 6 --launch-count 1 \
     -o gpurun_out/prof_bn_$T -f python tools/bn_bench.py --only $L > gpurun_out/ncu_bn_$T.log 2>&1
  tail -2 gpurun_out/ncu_bn_$T.log
done
