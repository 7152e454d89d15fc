#!/bin/bash
set -u
TAG=${1:-r2an}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tree_step -s 402 -c 1 -f -o /tmp/${TAG}_tree \
  python tools/bench3d.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_tree_ncu.log 2>&1
python tools/summarize_ncu.py /tmp/${TAG}_tree.ncu-rep > gpurun_out/${TAG}_tree.txt 2>&1
python tools/summarize_ncu.py /tmp/${TAG}_tree.ncu-rep --traffic > gpurun_out/${TAG}_tree_traffic.txt 2>&1
ncu -i /tmp/${TAG}_tree.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_tree_source.csv.gz
head -30 gpurun_out/${TAG}_tree.txt
