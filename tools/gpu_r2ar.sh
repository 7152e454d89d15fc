#!/bin/bash
# OSC-action rollout after tier 2: tier occupancy, and the lock-step CTA width A/B again (w1 / w3 / base = 7 warps)
set -u
TAG=${1:-r2ar}
mkdir -p gpurun_out
timeout 600 python tools/diag_tiers.py --collects 12 2>&1 | tail -14 | tee gpurun_out/${TAG}_tiers.txt
for lib in base w1 w3; do
  l=cassierl_b200/lib/libcassie2d.so; [ $lib != base ] && l=cassierl_b200/lib/libcassie2d_${lib}.so
  [ -f $l ] || continue
  CASSIE2D_LIB=$PWD/$l timeout 300 python tools/bench_rollout.py --mode OSC --T 20 --reps 7 2>&1 | tail -1 > gpurun_out/${TAG}_rollout_osc_${lib}.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_rollout_osc_${lib}.json')); print('$lib OSC rollout env-steps/s %.4g collect_ms %.2f' % (d['env_steps_per_s'], d['collect_ms']))"
  CASSIE2D_LIB=$PWD/$l timeout 300 python tools/bench_rollout.py --mode OSC --T 20 --reps 5 --envs 65536 2>&1 | tail -1 > gpurun_out/${TAG}_rollout_osc64k_${lib}.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_rollout_osc64k_${lib}.json')); print('$lib OSC rollout 65536 envs env-steps/s %.4g collect_ms %.2f' % (d['env_steps_per_s'], d['collect_ms']))"
  CASSIE2D_LIB=$PWD/$l timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${TAG}_bench_${lib}.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench_${lib}.json')); print('$lib squat_osc value %.4g ms %.4f' % (d['value'], d['ms_per_step']))"
done 2>&1 | tee gpurun_out/${TAG}_width_ab.txt
