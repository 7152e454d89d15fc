#!/bin/bash
# steady-state capture of one workload: skips the pre-advance launches and the warm-up, captures the first timed step
set -u
TAG=$1; wl=$2; kr=$3
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kr -s 101 -c 1 -f -o /tmp/${TAG}_${wl} \
  python bench.py --workload $wl --steps 2 --warmup 1 --preadvance 1000 --no-cpu-baseline > gpurun_out/${TAG}_${wl}_ncu.log 2>&1
python tools/summarize_ncu.py /tmp/${TAG}_${wl}.ncu-rep > gpurun_out/${TAG}_${wl}.txt 2>&1
ncu -i /tmp/${TAG}_${wl}.ncu-rep --page raw --csv > gpurun_out/${TAG}_${wl}_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_${wl}.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_${wl}_source.csv.gz
head -32 gpurun_out/${TAG}_${wl}.txt
