#!/bin/bash
set -u
TAG=${1:-r2j}
mkdir -p gpurun_out
echo "== pytest -m gpu (quad engine)"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.txt
for wl in ${WLS:-squat_osc squat_jacobian pd_env}; do
  echo "== bench quad $wl"; timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_${wl}.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('   value %.4g  e2e %.4g  frac %.4f  ms %.4f  rows %.2f->%.2f max %d qp %.2f' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['ms_per_step'], d['stats']['first_timed_step']['rows_mean'], d['stats']['last_step']['rows_mean'], d['stats']['last_step']['rows_max'], d['stats']['last_step']['qp_iters_mean']))
except Exception as e: print('   parse failed', e)
"
done
bash tools/gpu_prof3.sh ${TAG} ${PROF_WL:-squat_osc} ${PROF_K:-k_qsquat} | head -30
