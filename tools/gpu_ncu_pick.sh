#!/bin/bash
# ncu --set full of chosen launches of one eager f16x2 step: tools/gpu_ncu_pick.sh <launch index> ...
mkdir -p gpurun_out
for i in "$@"; do
  ncu --set full --clock-control none --import-source on --profile-from-start off --launch-skip $i --launch-count 1 -f \
      -o gpurun_out/prof_f16x2_l$i python tools/profile_step.py --steps 1 --precision f16x2 > gpurun_out/ncu_l$i.log 2>&1
  tail -2 gpurun_out/ncu_l$i.log
done
ls -la gpurun_out/*.ncu-rep | tail -5
