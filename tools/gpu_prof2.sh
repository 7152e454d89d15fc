#!/bin/bash
# ncu captures: tools/gpu_prof2.sh TAG "workload:kernel_regex:preadvance ..."   (CASSIE_ENGINE from the environment).
# The reports are summarised ON the box (raw page, source page as csv.gz) and deleted: gpurun_out/ is capped at 64 MiB.
set -u
TAG=${1:-prof}; shift
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=: read wl kr pre <<< "$spec"
  echo "== bench $wl pre=$pre"; timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --preadvance $pre --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_${wl}.json | cut -c1-300
  echo "== ncu $wl $kr"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kr -s ${SKIP:-3} -c 1 -f -o /tmp/${TAG}_${wl} \
    python bench.py --workload $wl --steps 2 --warmup 1 --preadvance $pre --no-cpu-baseline > gpurun_out/${TAG}_${wl}_ncu.log 2>&1
  tail -2 gpurun_out/${TAG}_${wl}_ncu.log | cut -c1-200
  python tools/summarize_ncu.py /tmp/${TAG}_${wl}.ncu-rep > gpurun_out/${TAG}_${wl}.txt 2>&1
  ncu -i /tmp/${TAG}_${wl}.ncu-rep --page raw --csv > gpurun_out/${TAG}_${wl}_raw.csv 2>/dev/null
  ncu -i /tmp/${TAG}_${wl}.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_${wl}_source.csv.gz
  head -40 gpurun_out/${TAG}_${wl}.txt
done
ls -la gpurun_out | tail -8
