#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench lines for every workload, ncu launch list and one
# full capture of the top kernel.  Everything lands in gpurun_out/.
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee gpurun_out/${TAG}_smoke.txt
echo "== bench (headline)"; timeout 600 python bench.py --steps 30 --warmup 5 2>&1 | tail -3 | tee gpurun_out/${TAG}_bench.json
for wl in ${WORKLOADS:-squat_jacobian torque_random pd_env}; do
  echo "== bench $wl"; timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/${TAG}_bench_${wl}.json
done
echo "== bench 131072 envs"; timeout 300 python bench.py --envs 131072 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/${TAG}_bench_131072.json
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 10 --warmup 2 2>&1 | tail -2 | tee gpurun_out/${TAG}_bench_reference.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1
tail -5 gpurun_out/${TAG}_launches.csv
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-k_squat} -s 2 -c 1 -f -o gpurun_out/${TAG}_prof \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -20
