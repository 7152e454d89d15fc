#!/bin/bash
# after tier 2: where do the envs of the OSC-action rollout sit, and does a larger batch hide the slow ones?
set -u
TAG=${1:-r2aq}
mkdir -p gpurun_out
timeout 600 python tools/diag_tiers.py --collects 12 2>&1 | tail -14 | tee gpurun_out/${TAG}_tiers.txt
for n in 16384 65536 131072 262144; do
  timeout 600 python tools/bench_rollout.py --mode OSC --T 20 --reps 5 --envs $n 2>&1 | tail -1 > gpurun_out/${TAG}_rollout_OSC_${n}.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_rollout_OSC_${n}.json')); print('rollout OSC envs $n env-steps/s %.4g collect_ms %.3f' % (d['env_steps_per_s'], d['collect_ms']), d['last_step_qp'])"
done 2>&1 | tee gpurun_out/${TAG}_envs_sweep.txt
