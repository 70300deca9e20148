#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bn_bench.py > gpurun_out/bn_new.log 2>&1; grep -v "^{" gpurun_out/bn_new.log | tail -20; tail -1 gpurun_out/bn_new.log | cut -c1-200
PPY_BN_REPLICAS=1 timeout 300 python tools/bn_bench.py > gpurun_out/bn_old.log 2>&1; grep -v "^{" gpurun_out/bn_old.log | tail -20; tail -1 gpurun_out/bn_old.log | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_train.py -q -p no:cacheprovider -x -k "bn or batchnorm or BatchNorm" 2>&1 | tail -3
