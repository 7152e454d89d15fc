#!/bin/bash
# Round-2 final verification: whole GPU suite, smoke, headline + reference arm, 3-D bench with CPU baseline, captures
set -u
TAG=${1:-r2w}
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/${TAG}_smoke.txt
echo "== headline"; timeout 600 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 > gpurun_out/${TAG}_bench.json; cut -c1-300 gpurun_out/${TAG}_bench.json
echo "== reference arm"; timeout 400 python bench.py --impl reference --steps 10 --warmup 2 2>&1 | tail -1 > gpurun_out/${TAG}_bench_reference.json; cut -c1-200 gpurun_out/${TAG}_bench_reference.json
echo "== bench3d (CPU baseline)"; timeout 900 python tools/bench3d.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/${TAG}_bench3d.json; cut -c1-200 gpurun_out/${TAG}_bench3d.json
for spec in "32 2" "16 2" "16 4" "8 4"; do
  set -- $spec
  CASSIE3D_TILES=$2 timeout 600 python tools/bench3d.py --lanes $1 --steps 10 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${TAG}_bench3d_l$1_t$2.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench3d_l$1_t$2.json')); print('lanes $1 tiles/CTA $2 (step barrier on) value %.4g e2e %.4g ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))" 2>&1 | tail -1
done | tee gpurun_out/${TAG}_lanes.txt
echo "== ncu tree (final)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tree_step -s 201 -c 1 -f -o /tmp/${TAG}_tree \
  python tools/bench3d.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_tree_ncu.log 2>&1
python tools/summarize_ncu.py /tmp/${TAG}_tree.ncu-rep > gpurun_out/${TAG}_tree.txt 2>&1
python tools/summarize_ncu.py /tmp/${TAG}_tree.ncu-rep --traffic > gpurun_out/${TAG}_tree_traffic.txt 2>&1
head -30 gpurun_out/${TAG}_tree.txt
echo "== ncu launch list of tools/bench3d.py"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file gpurun_out/${TAG}_launches3d.csv \
  python tools/bench3d.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch3d.log 2>&1
tail -2 gpurun_out/${TAG}_launches3d.csv | cut -c1-200
