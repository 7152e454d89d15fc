#!/bin/bash
# bench the experimental library variants (cassierl_b200/build.py CASSIE2D_VARIANT) side by side:
#   tools/gpu_variants.sh [-w "workloads"] variant...
mkdir -p gpurun_out
WL="pd_env squat_jacobian squat_osc"
if [ "$1" = "-w" ]; then WL="$2"; shift 2; fi
for v in "" "$@"; do
  lib=cassierl_b200/lib/libcassie2d${v:+_$v}.so
  for wl in $WL; do
    echo -n "variant=${v:-base} $wl: "
    CASSIE2D_LIB=$PWD/$lib python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('%.3e env-steps/s  %.3f ms' % (d['value'], d['ms_per_step']))"
  done
done 2>&1 | tee gpurun_out/variants.txt
