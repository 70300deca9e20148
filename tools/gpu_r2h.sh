#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q -p no:cacheprovider -x -k "pack_weight_dgrad or bn_act_backward" 2>&1 | tail -15 | cut -c1-250
