#!/bin/bash
# Evidence run of round 2 (under gpurun from the repo root): ncu launch list of one eager step of the HEADLINE (f16x2) workload
# with DRAM bytes / tensor-pipe / occupancy per launch, and --set full captures of three representative launches
# (an HBM-bound 1x1 + residual layer, a tensor-bound 3x3 layer, the whole-layer DCN kernel).  Outputs under gpurun_out/.
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__registers_per_thread
ncu --metrics $M --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches_f16x2.csv \
    python tools/profile_step.py --steps 1 --precision f16x2 > gpurun_out/profile_step.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_f16x2.csv gpurun_out/plan_names.txt gpurun_out/conv_family_traffic_f16x2.json > gpurun_out/launch_list_f16x2.md 2>&1
tail -14 gpurun_out/launch_list_f16x2.md
for i in "$@"; do
  ncu --set full --clock-control none --import-source on --profile-from-start off --launch-skip $i --launch-count 1 -f \
      -o gpurun_out/prof_f16x2_l$i python tools/profile_step.py --steps 1 --precision f16x2 > gpurun_out/ncu_l$i.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
