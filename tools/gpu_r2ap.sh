#!/bin/bash
# Tier 2 of the quad engine (four joint-limit slots per leg): GPU parity, headline no-regression, OSC-action rollout
set -u
TAG=${1:-r2ap}
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.txt
echo "== headline"; timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/${TAG}_bench.json
python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print('headline value %.4g e2e %.4g frac %.4f ms %.4f' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['ms_per_step']))" | tee gpurun_out/${TAG}_summary.txt
echo "== rollouts"
for e in quad thread; do
  for mode in OSC PD; do
    CASSIE_ENGINE=$e timeout 600 python tools/bench_rollout.py --mode $mode --T 20 --reps 7 2>&1 | tail -1 > gpurun_out/${TAG}_rollout_${mode}_${e}.json
    python -c "import json; d=json.load(open('gpurun_out/${TAG}_rollout_${mode}_${e}.json')); print('rollout $mode engine $e env-steps/s %.4g collect_ms %.3f' % (d['env_steps_per_s'], d['collect_ms']), {k: d[k] for k in d if 'qp' in k or 'rows' in k or 'stat' in k})"
  done
done 2>&1 | tee -a gpurun_out/${TAG}_summary.txt
echo "== OSC diag"
CASSIE_ENGINE=quad timeout 900 python tools/diag_osc_rollout.py 60 2>&1 | tail -16 | tee gpurun_out/${TAG}_osc_diag.txt
