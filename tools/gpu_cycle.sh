#!/bin/bash
# one GPU round-trip: gpu tests, bench, per-layer roofline table (run under gpurun from the repo root)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1
python tools/roofline_table.py gpurun_out/per_op_ms.json 90 > gpurun_out/roofline_table.txt 2>&1
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench.log').read().strip().split('\n')[-1])
print('img/s %.1f  ms %.3f  e2e %.1f  frac %.3f  conv_ms %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step']))
PY
