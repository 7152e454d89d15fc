#!/bin/bash
# round-2 first GPU call: the tightened parity tests + steady-state baseline numbers of the thread-per-env kernels
set -u
TAG=${1:-r2a}
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | tail -40 | tee gpurun_out/${TAG}_pytest.txt
echo "== bench (headline, steady state)"; timeout 600 python bench.py --steps 30 --warmup 5 2>&1 | tail -2 | tee gpurun_out/${TAG}_bench.json
for wl in squat_jacobian torque_random pd_env; do
  echo "== bench $wl"; timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_${wl}.json | cut -c1-600
done
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 10 --warmup 2 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_reference.json
