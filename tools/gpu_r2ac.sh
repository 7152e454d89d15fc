#!/bin/bash
set -u
TAG=${1:-r2ac}
mkdir -p gpurun_out
echo "== pytest tree"; timeout 1200 python -m pytest tests/test_gpu_tree.py -m gpu -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest_tree.txt
for spec in "7 1" "7 -1" "4 -1" "7 1" "7 -1"; do
  set -- $spec
  CASSIE3D_TILES=$1 CASSIE3D_STEP_BARRIER=$2 timeout 600 python tools/bench3d.py --lanes 32 --steps 10 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${TAG}_bench3d_t$1_b$2.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench3d_t$1_b$2.json')); print('lanes 32 tiles/CTA $1 barrier mode $2 value %.4g e2e %.4g ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))" 2>&1 | tail -1
done | tee gpurun_out/${TAG}_sweep.txt
