#!/bin/bash
# 3-D engine v6 (register-resident PGS for every tile width): tree tests + lanes x tiles sweep
set -u
TAG=${1:-r2y}
mkdir -p gpurun_out
echo "== pytest tree"; timeout 1200 python -m pytest tests/test_gpu_tree.py -m gpu -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest_tree.txt
for spec in "32 2" "32 7" "16 2" "16 4" "16 6" "16 14" "8 4" "8 8" "8 12"; do
  set -- $spec
  CASSIE3D_TILES=$2 timeout 600 python tools/bench3d.py --lanes $1 --steps 10 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${TAG}_bench3d_l$1_t$2.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench3d_l$1_t$2.json')); print('lanes $1 tiles/CTA $2 value %.4g e2e %.4g ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))" 2>&1 | tail -1
done | tee gpurun_out/${TAG}_sweep.txt
