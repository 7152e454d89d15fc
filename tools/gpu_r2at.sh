#!/bin/bash
# 8x B200: config 5 with the reference's default (OSC) action space after tier 2 / per-CTA tiers, and the headline again
set -u
TAG=${1:-r2at}
N=${2:-8}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 "$@"; }
run tools/bench_rollout.py --mode OSC --T 20 --reps 5 2>&1 | grep '^{' | tail -1 > gpurun_out/${TAG}_rollout_OSC_n${N}.json
python -c "import json; d=json.load(open('gpurun_out/${TAG}_rollout_OSC_n${N}.json')); print('N=$N rollout OSC env-steps/s %.4g collect_ms %.2f with overlapped gather %.4g' % (d['env_steps_per_s'], d['collect_ms'], d['env_steps_per_s_with_overlapped_gather'] or 0), {k: d['gather'].get(k) for k in ('gather_ms', 'bytes_per_rank', 'collective')})" | tee gpurun_out/${TAG}_summary.txt
[ -n "${SKIP64K:-}" ] || run tools/bench_rollout.py --mode OSC --T 20 --reps 5 --envs 65536 2>&1 | grep '^{' | tail -1 > gpurun_out/${TAG}_rollout_OSC64k_n${N}.json
[ -n "${SKIP64K:-}" ] || python -c "import json; d=json.load(open('gpurun_out/${TAG}_rollout_OSC64k_n${N}.json')); print('N=$N rollout OSC 65536 envs per GPU env-steps/s %.4g collect_ms %.2f' % (d['env_steps_per_s'], d['collect_ms']))" | tee -a gpurun_out/${TAG}_summary.txt
run bench.py --gpus $N --steps 10 --warmup 3 2>&1 | grep '^{' | tail -1 > gpurun_out/${TAG}_bench_n${N}.json
python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench_n${N}.json')); print('N=$N headline value %.4g e2e %.4g ms %.4f' % (d['value'], d['e2e']['value'], d['ms_per_step']))" | tee -a gpurun_out/${TAG}_summary.txt
