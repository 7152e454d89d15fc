#!/bin/bash
# final-tree verification of the session that added tier 2 and the per-CTA tier vote
set -u
TAG=${1:-r2fin3}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee gpurun_out/${TAG}_smoke.txt
echo "== reference arm"; timeout 400 python bench.py --impl reference 2>&1 | tail -1 > gpurun_out/${TAG}_bench_reference.json
echo "== bench (headline)"; timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/${TAG}_bench.json
python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); r=json.load(open('gpurun_out/${TAG}_bench_reference.json')); print('headline value %.4g e2e %.4g frac %.4f ms %.4f launches %s | reference arm %.4g (%s)' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['ms_per_step'], d.get('gpu_launches'), r['value'], r['metric'] == d['metric']))" | tee gpurun_out/${TAG}_summary.txt
for wl in squat_jacobian pd_env; do
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${TAG}_bench_${wl}.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench_${wl}.json')); print('$wl value %.4g ms %.4f' % (d['value'], d['ms_per_step']))"
done 2>&1 | tee -a gpurun_out/${TAG}_summary.txt
for mode in OSC PD; do
  timeout 600 python tools/bench_rollout.py --mode $mode --T 20 --reps 7 2>&1 | tail -1 > gpurun_out/${TAG}_rollout_${mode}.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_rollout_${mode}.json')); print('rollout $mode (default engine) env-steps/s %.4g collect_ms %.3f' % (d['env_steps_per_s'], d['collect_ms']), d['last_step_qp'])"
done 2>&1 | tee -a gpurun_out/${TAG}_summary.txt
timeout 300 python tools/bench_rollout.py --mode OSC --T 20 --reps 5 --envs 65536 2>&1 | tail -1 > gpurun_out/${TAG}_rollout_osc64k.json
python -c "import json; d=json.load(open('gpurun_out/${TAG}_rollout_osc64k.json')); print('OSC rollout 65536 envs env-steps/s %.4g collect_ms %.2f' % (d['env_steps_per_s'], d['collect_ms']))" | tee -a gpurun_out/${TAG}_summary.txt
timeout 600 python tools/bench3d.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${TAG}_bench3d.json
python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench3d.json')); print('3-D value %.4g e2e %.4g' % (d['value'], d['e2e']['value']))" | tee -a gpurun_out/${TAG}_summary.txt
echo "== ncu launch list (OSC rollout)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches_rollout.csv \
  python tools/bench_rollout.py --mode OSC --T 20 --reps 2 > gpurun_out/${TAG}_ncu_launch.log 2>&1
grep -c k_qrollout gpurun_out/${TAG}_launches_rollout.csv
