#!/bin/bash
# ncu captures of the step kernels: tools/gpu_prof.sh TAG "workload:kernel_regex ..."
set -u
TAG=${1:-prof}; shift
mkdir -p gpurun_out
for spec in "$@"; do
  wl=${spec%%:*}; kr=${spec##*:}
  echo "== bench $wl"; timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_${wl}.json | cut -c1-400
  echo "== ncu $wl $kr"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kr -s 2 -c 1 -f -o gpurun_out/${TAG}_${wl} \
    python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_${wl}_ncu.log 2>&1
  tail -2 gpurun_out/${TAG}_${wl}_ncu.log | cut -c1-200
done
ls -la gpurun_out | tail -8
