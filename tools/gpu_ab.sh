#!/bin/bash
# A/B of library variants: tools/gpu_ab.sh TAG "variant ..." "workload ..."   (variant "base" = libcassie2d.so)
set -u
TAG=$1; VARS=$2; WLS=$3
mkdir -p gpurun_out
for v in $VARS; do
  lib=cassierl_b200/lib/libcassie2d.so; [ "$v" != base ] && lib=cassierl_b200/lib/libcassie2d_$v.so
  for wl in $WLS; do
    CASSIE2D_LIB=$PWD/$lib timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${TAG}_${v}_${wl}.json
    python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_${v}_${wl}.json')); print('%-6s %-15s value %.4g  e2e %.4g  frac %.4f  ms %.4f' % ('$v','$wl',d['value'], d['e2e']['value'], d['roofline']['frac'], d['ms_per_step']))
except Exception as e: print('$v $wl failed', e, open('gpurun_out/${TAG}_${v}_${wl}.json').read()[-300:])
PY
  done
done
