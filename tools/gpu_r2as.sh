#!/bin/bash
# constraint tier chosen per CTA (cta_any) instead of per warp: parity, headline, rollouts
set -u
TAG=${1:-r2as}
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.txt
echo "== headline"; timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/${TAG}_bench.json
python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print('headline value %.4g e2e %.4g frac %.4f ms %.4f' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['ms_per_step']))" | tee gpurun_out/${TAG}_summary.txt
for mode in OSC PD; do
  CASSIE_ENGINE=quad timeout 600 python tools/bench_rollout.py --mode $mode --T 20 --reps 7 2>&1 | tail -1 > gpurun_out/${TAG}_rollout_${mode}_quad.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_rollout_${mode}_quad.json')); print('rollout $mode engine quad env-steps/s %.4g collect_ms %.3f' % (d['env_steps_per_s'], d['collect_ms']))"
done 2>&1 | tee -a gpurun_out/${TAG}_summary.txt
timeout 300 python tools/bench_rollout.py --mode OSC --T 20 --reps 5 --envs 65536 2>&1 | tail -1 > gpurun_out/${TAG}_rollout_osc64k.json
python -c "import json; d=json.load(open('gpurun_out/${TAG}_rollout_osc64k.json')); print('OSC rollout 65536 envs env-steps/s %.4g collect_ms %.2f' % (d['env_steps_per_s'], d['collect_ms']))" | tee -a gpurun_out/${TAG}_summary.txt
CASSIE_ENGINE=quad timeout 900 python tools/diag_osc_rollout.py 60 2>&1 | tail -16 | tee gpurun_out/${TAG}_osc_diag.txt
