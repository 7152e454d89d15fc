#!/bin/bash
set -u
TAG=${1:-r2aa}
mkdir -p gpurun_out
echo "== pytest tree"; timeout 1200 python -m pytest tests/test_gpu_tree.py -m gpu -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest_tree.txt
for spec in "32 2" "32 7" "16 2"; do
  set -- $spec
  CASSIE3D_TILES=$2 timeout 600 python tools/bench3d.py --lanes $1 --steps 10 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${TAG}_bench3d_l$1_t$2.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench3d_l$1_t$2.json')); print('lanes $1 tiles/CTA $2 value %.4g e2e %.4g ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))" 2>&1 | tail -1
done | tee gpurun_out/${TAG}_sweep.txt
echo "== ncu"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tree_step -s 201 -c 1 -f -o /tmp/${TAG}_tree \
  python tools/bench3d.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_tree_ncu.log 2>&1
python tools/summarize_ncu.py /tmp/${TAG}_tree.ncu-rep > gpurun_out/${TAG}_tree.txt 2>&1
ncu -i /tmp/${TAG}_tree.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_tree_source.csv.gz
head -30 gpurun_out/${TAG}_tree.txt
