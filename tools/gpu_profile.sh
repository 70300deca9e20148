#!/bin/bash
# Evidence run (under gpurun from the repo root): ncu launch list of one eager step of the bench workload with DRAM bytes,
# tensor-pipe and occupancy metrics per launch, plus the bench line and the per-layer table.  Outputs under gpurun_out/.
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__registers_per_thread
ncu --metrics $M --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches.csv \
    python tools/profile_step.py --steps 1 > gpurun_out/profile_step.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv gpurun_out/plan_names.txt > gpurun_out/launch_list.md 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1
python tools/roofline_table.py gpurun_out/per_op_ms.json 200 > gpurun_out/roofline_table.txt 2>&1
tail -12 gpurun_out/launch_list.md; tail -1 gpurun_out/bench.log | cut -c1-400
