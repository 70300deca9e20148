#!/bin/bash
# ncu --set full captures of representative conv launches of one eager step (under gpurun; launch indices = rows of the launch list)
mkdir -p gpurun_out
for i in "$@"; do
  ncu --set full --clock-control none --import-source on --profile-from-start off --launch-skip $i --launch-count 1 -f \
      -o gpurun_out/prof_l$i python tools/profile_step.py --steps 1 > gpurun_out/ncu_l$i.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
