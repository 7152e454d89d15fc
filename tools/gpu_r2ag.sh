#!/bin/bash
# configs[3] on 8 GPUs with the final 3-D build (65536 envs), + the headline at N = 8 once more
set -u
TAG=${1:-r2ag}
mkdir -p gpurun_out
run() { local n=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) "$@"; }
for n in 8; do
  run $n tools/bench3d.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 > gpurun_out/${TAG}_bench3d_n${n}.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench3d_n${n}.json')); print('bench3d N=%d envs total %d value %.4g e2e %.4g ms %.3f resets %d dropped %d non-finite %d' % (d['n_gpus'], d['config']['envs_total'], d['value'], d['e2e']['value'], d['ms_per_step'], d['stats']['auto_resets_in_timed_region'], d['stats']['contacts_dropped'], d['stats']['non_finite_envs']))"
  run $n bench.py --gpus $n --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 > gpurun_out/${TAG}_bench_n${n}.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench_n${n}.json')); print('bench weak N=%d value %.4g e2e %.4g ms %.4f' % (d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step']))"
done | tee gpurun_out/${TAG}_scaling.txt
