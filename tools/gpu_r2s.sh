#!/bin/bash
# Multi-GPU evidence on ONE box with G GPUs (gpurun --gpus G): NCCL gather test, config 5 rollouts, configs[3] (3-D),
# weak and strong scaling of the headline.  usage: tools/gpu_r2s.sh TAG "1 2" (GPU counts to run)
set -u
TAG=${1:-r2s}; NS=${2:-"1 2"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | tee gpurun_out/${TAG}_gpus.txt
echo "== NCCL gather test"; timeout 900 python -m pytest tests/test_gpu_nccl.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest_nccl.txt
run() {  # run N script args...
  local n=$1; shift
  if [ "$n" = "1" ]; then timeout 900 python "$@"; else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) "$@"; fi
}
for n in $NS; do
  run $n tools/bench_rollout.py --mode PD --T 20 2>&1 | grep '^{' | tail -1 > gpurun_out/${TAG}_rollout_pd_n${n}.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_rollout_pd_n${n}.json')); g=d['gather']
    print('rollout PD  N=%d  env-steps/s %.4g  collect %.2f ms  gather %s ms (%s B/rank in, %s GB/s)  overlapped %s ms  -> %s env-steps/s with the gather hidden' % (d['n_gpus'], d['env_steps_per_s'], d['collect_ms'], g.get('gather_ms'), g.get('bytes_per_rank'), g.get('gather_GBps_per_rank_in'), g.get('collect_with_overlapped_gather_ms'), d.get('env_steps_per_s_with_overlapped_gather')))
except Exception as e: print('rollout N=$n failed', e)
PY
  run $n tools/bench3d.py --steps 10 --warmup 2 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 > gpurun_out/${TAG}_bench3d_n${n}.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench3d_n${n}.json')); print('bench3d     N=%d  envs total %d  value %.4g  e2e %.4g  ms %.3f' % (d['n_gpus'], d['config']['envs_total'], d['value'], d['e2e']['value'], d['ms_per_step']))" 2>&1 | tail -1
  run $n bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 > gpurun_out/${TAG}_bench_weak_n${n}.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench_weak_n${n}.json')); print('bench weak  N=%d  value %.4g  e2e %.4g  ms %.4f' % (d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step']))" 2>&1 | tail -1
  for tot in 16384 131072; do
    run $n bench.py --gpus $n --scaling strong --envs $tot --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 > gpurun_out/${TAG}_bench_strong${tot}_n${n}.json
    python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench_strong${tot}_n${n}.json')); print('bench strong N=%d  envs total $tot  value %.4g  e2e %.4g  ms %.4f' % (d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step']))" 2>&1 | tail -1
  done
done | tee gpurun_out/${TAG}_scaling.txt
