#!/bin/bash
# 3-D tree engine after the PGS specialisation: full GPU suite, lanes sweep, CPU baseline, ncu capture
set -u
TAG=${1:-r2t}
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.txt
for lanes in 32 16 8; do
  timeout 600 python tools/bench3d.py --lanes $lanes --steps 10 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${TAG}_bench3d_l${lanes}.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench3d_l${lanes}.json')); print('lanes $lanes value %.4g e2e %.4g frac %.4f ms %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['ms_per_step']), d['stats'])" 2>&1 | tail -1
done | tee gpurun_out/${TAG}_lanes.txt
for t in 2 3 4; do
  CASSIE3D_TILES=$t timeout 600 python tools/bench3d.py --lanes 32 --steps 10 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${TAG}_bench3d_t${t}.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench3d_t${t}.json')); print('lanes 32 tiles/CTA $t value %.4g ms %.3f' % (d['value'], d['ms_per_step']))" 2>&1 | tail -1
done | tee -a gpurun_out/${TAG}_lanes.txt
echo "== bench3d with the CPU baseline (default lanes)"
timeout 900 python tools/bench3d.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/${TAG}_bench3d.json; cut -c1-600 gpurun_out/${TAG}_bench3d.json
echo "== ncu"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tree_step -s 201 -c 1 -f -o /tmp/${TAG}_tree \
  python tools/bench3d.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_tree_ncu.log 2>&1
python tools/summarize_ncu.py /tmp/${TAG}_tree.ncu-rep > gpurun_out/${TAG}_tree.txt 2>&1
python tools/summarize_ncu.py /tmp/${TAG}_tree.ncu-rep --traffic > gpurun_out/${TAG}_tree_traffic.txt 2>&1
ncu -i /tmp/${TAG}_tree.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_tree_source.csv.gz
head -34 gpurun_out/${TAG}_tree.txt
