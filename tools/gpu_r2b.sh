#!/bin/bash
# quad engine: parity tests + A/B bench against the thread engine on all four workloads
set -u
TAG=${1:-r2b}
mkdir -p gpurun_out
echo "== pytest -m gpu (quad engine)"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 | tee gpurun_out/${TAG}_pytest.txt
for eng in quad thread; do
for wl in squat_osc squat_jacobian torque_random pd_env; do
  echo "== bench $eng $wl"; CASSIE_ENGINE=$eng timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_${eng}_${wl}.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('   value %.4g  e2e %.4g  frac %.4f  ms %.4f  rows %s->%s sweeps %.1f qp %.2f' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['ms_per_step'], d['stats']['first_timed_step']['rows_mean'], d['stats']['last_step']['rows_mean'], d['stats']['last_step']['pgs_sweeps_mean'], d['stats']['last_step']['qp_iters_mean']))
except Exception as e: print('   parse failed', e)
"
done
done
