#!/bin/bash
set -u
TAG=${1:-r2ai}
mkdir -p gpurun_out
echo "== pytest tree"; timeout 1200 python -m pytest tests/test_gpu_tree.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest_tree.txt
for srt in 1 0 1 0; do
  CASSIE3D_SORT=$srt timeout 600 python tools/bench3d.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${TAG}_bench3d_s$srt.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench3d_s$srt.json')); print('binning $srt value %.4g e2e %.4g ms %.3f launches %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches']))"
done | tee gpurun_out/${TAG}_runs.txt
