#!/bin/bash
# Round-2 evidence in ONE gpurun call: GPU parity tests, smoke, engine x workload matrix, headline + reference arm,
# rollout bench (PD / OSC action spaces, both engines), ncu launch list and one steady-state full capture of the quad
# OSC squat kernel (summarised on the box; the reports are too big for gpurun_out/).
set -u
TAG=${1:-r2p}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/${TAG}_smoke.txt
echo "== headline"; timeout 600 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 > gpurun_out/${TAG}_bench.json; cut -c1-400 gpurun_out/${TAG}_bench.json
echo "== reference arm"; timeout 400 python bench.py --impl reference --steps 10 --warmup 2 2>&1 | tail -1 > gpurun_out/${TAG}_bench_reference.json; cut -c1-300 gpurun_out/${TAG}_bench_reference.json
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
done | tee gpurun_out/${TAG}_engines.txt
echo "== strong-scaling points on one GPU (envs per GPU of a 16384-env job on 2/4/8 GPUs)"
for n in 8192 4096 2048 131072; do
  timeout 300 python bench.py --envs $n --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${TAG}_bench_envs${n}.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench_envs${n}.json')); print('envs %6d  value %.4g  e2e %.4g  ms %.4f' % ($n, d['value'], d['e2e']['value'], d['ms_per_step']))"
done | tee gpurun_out/${TAG}_envs.txt
echo "== rollout"
for eng in quad thread; do for mode in PD OSC; do
  CASSIE_ENGINE=$eng timeout 300 python tools/bench_rollout.py --mode $mode --T 20 2>&1 | tail -1 > gpurun_out/${TAG}_rollout_${eng}_${mode}.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_rollout_${eng}_${mode}.json')); print('$eng $mode env-steps/s %.4g collect_ms %.3f post_ms %.3f qp %s' % (d['env_steps_per_s'], d['collect_ms'], d['returns_baseline_advantages_ms'], d['last_step_qp']))"
done; done | tee gpurun_out/${TAG}_rollout.txt
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1
tail -4 gpurun_out/${TAG}_launches.csv | cut -c1-250
echo "== ncu full (steady state: skip pre-advance + warm-up launches)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_qsquat -s 101 -c 1 -f -o /tmp/${TAG}_squat_osc \
  python bench.py --workload squat_osc --steps 2 --warmup 1 --preadvance 1000 --no-cpu-baseline > gpurun_out/${TAG}_squat_osc_ncu.log 2>&1
python tools/summarize_ncu.py /tmp/${TAG}_squat_osc.ncu-rep > gpurun_out/${TAG}_squat_osc.txt 2>&1
python tools/summarize_ncu.py /tmp/${TAG}_squat_osc.ncu-rep --traffic > gpurun_out/${TAG}_squat_osc_traffic.txt 2>&1
ncu -i /tmp/${TAG}_squat_osc.ncu-rep --page raw --csv > gpurun_out/${TAG}_squat_osc_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_squat_osc.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_squat_osc_source.csv.gz
head -40 gpurun_out/${TAG}_squat_osc.txt
ls -la gpurun_out | tail -12
