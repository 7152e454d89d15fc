#!/bin/bash
# 3-D tree engine on the GPU: parity tests, lanes sweep of tools/bench3d.py, one ncu capture
set -u
TAG=${1:-r2q}
mkdir -p gpurun_out
echo "== pytest -m gpu (whole suite, no -x: every failure is listed)"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/${TAG}_smoke.txt
for lanes in 32 16 8; do
  timeout 600 python tools/bench3d.py --lanes $lanes --steps 10 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${TAG}_bench3d_l${lanes}.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench3d_l${lanes}.json')); print('lanes $lanes value %.4g e2e %.4g frac %.4f ms %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['ms_per_step']), d['stats'])" 2>&1 | tail -1
done | tee gpurun_out/${TAG}_lanes.txt
echo "== ncu"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tree_step -s 201 -c 1 -f -o /tmp/${TAG}_tree \
  python tools/bench3d.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_tree_ncu.log 2>&1
python tools/summarize_ncu.py /tmp/${TAG}_tree.ncu-rep > gpurun_out/${TAG}_tree.txt 2>&1
ncu -i /tmp/${TAG}_tree.ncu-rep --page raw --csv > gpurun_out/${TAG}_tree_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_tree.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_tree_source.csv.gz
head -36 gpurun_out/${TAG}_tree.txt
echo "== tensor-core MLP A/B (thread-engine rollout kernel, PD action space)"
for v in scalar tc scalar tc; do
  CASSIE_MLP=$v timeout 300 python tools/bench_rollout.py --mode PD --T 20 --reps 7 2>&1 | tail -1 > gpurun_out/${TAG}_rollout_mlp_${v}.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_rollout_mlp_${v}.json')); print('mlp $v env-steps/s %.4g collect_ms %.3f' % (d['env_steps_per_s'], d['collect_ms']))"
done | tee gpurun_out/${TAG}_mlp_ab.txt
for v in scalar tc; do
  CASSIE_MLP=$v timeout 300 python tools/bench_rollout.py --mode PD --task imitate --T 20 --reps 7 2>&1 | tail -1 > gpurun_out/${TAG}_rollout_mlp_imitate_${v}.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_rollout_mlp_imitate_${v}.json')); print('mlp $v (imitate, 26 inputs) env-steps/s %.4g collect_ms %.3f' % (d['env_steps_per_s'], d['collect_ms']))"
done | tee -a gpurun_out/${TAG}_mlp_ab.txt
echo "== OSC rollout diagnosis"
bash tools/gpu_r2r.sh ${TAG}
ls gpurun_out | wc -l
