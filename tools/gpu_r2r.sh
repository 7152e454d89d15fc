#!/bin/bash
# OSC-action rollout diagnosis + A/B of the lock-step CTA width (variant w1: one warp per CTA for the controller modes)
set -u
TAG=${1:-r2r}
mkdir -p gpurun_out
for eng in quad thread; do
  CASSIE_ENGINE=$eng timeout 600 python tools/diag_osc_rollout.py 60 2>&1 | tail -16 | tee gpurun_out/${TAG}_diag_${eng}.txt
done
if [ -f cassierl_b200/lib/libcassie2d_w1.so ]; then
  CASSIE2D_LIB=$PWD/cassierl_b200/lib/libcassie2d_w1.so CASSIE_ENGINE=quad timeout 600 python tools/diag_osc_rollout.py 60 2>&1 | tail -4 | tee gpurun_out/${TAG}_diag_quad_w1.txt
  for lib in base w1; do
    l=cassierl_b200/lib/libcassie2d.so; [ $lib = w1 ] && l=cassierl_b200/lib/libcassie2d_w1.so
    CASSIE2D_LIB=$PWD/$l timeout 300 python tools/bench_rollout.py --mode OSC --T 20 2>&1 | tail -1 > gpurun_out/${TAG}_rollout_osc_${lib}.json
    python -c "import json; d=json.load(open('gpurun_out/${TAG}_rollout_osc_${lib}.json')); print('$lib OSC rollout env-steps/s %.4g collect_ms %.2f' % (d['env_steps_per_s'], d['collect_ms']), d['last_step_qp'])"
    CASSIE2D_LIB=$PWD/$l timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${TAG}_bench_${lib}.json
    python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench_${lib}.json')); print('$lib squat_osc value %.4g ms %.4f' % (d['value'], d['ms_per_step']))"
  done | tee gpurun_out/${TAG}_w1_ab.txt
fi
