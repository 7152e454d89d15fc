#!/bin/bash
set -u
TAG=${1:-r2n}
mkdir -p gpurun_out
echo "== pytest -m gpu (quad engine)"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.txt
for eng in quad thread; do
for wl in squat_osc squat_jacobian pd_env torque_random; do
  CASSIE_ENGINE=$eng timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${TAG}_bench_${eng}_${wl}.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_bench_${eng}_${wl}.json')); print('%-7s %-15s value %.4g  e2e %.4g  frac %.4f  ms %.4f  rows %.2f sweeps %.1f' % ('$eng','$wl',d['value'], d['e2e']['value'], d['roofline']['frac'], d['ms_per_step'], d['stats']['last_step']['rows_mean'], d['stats']['last_step']['pgs_sweeps_mean']))
except Exception as e: print('$eng $wl failed', e)
PY
done
done
