#!/bin/bash
set -u
TAG=${1:-r2ao}
mkdir -p gpurun_out
for spec in "0 0" "0 1" "5 0" "5 1" "7 0" "3 0"; do
  set -- $spec
  CASSIE3D_TILES=$1 CASSIE3D_SORT=$2 timeout 600 python tools/bench3d.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${TAG}_bench3d_t$1_s$2.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench3d_t$1_s$2.json')); print('tiles/CTA $1 (0 = default 10) binning $2 value %.4g e2e %.4g ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done | tee gpurun_out/${TAG}_sweep.txt
