#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -q -p no:cacheprovider -k "nms" > gpurun_out/t_nms.log 2>&1
tail -4 gpurun_out/t_nms.log
