#!/bin/bash
set -u
TAG=${1:-r2al}
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.txt
for tp in 1 0 1 0; do
  CASSIE3D_TWO_PASS=$tp timeout 600 python tools/bench3d.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${TAG}_bench3d_tp$tp.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench3d_tp$tp.json')); print('two_pass $tp value %.4g e2e %.4g ms %.3f launches %d dropped %d resets %d smem/env %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['stats']['contacts_dropped'], d['stats']['auto_resets_in_timed_region'], d['stats']['smem_bytes_per_env']))"
done | tee gpurun_out/${TAG}_runs.txt
echo "== soak"; timeout 600 python tools/soak3d.py 1000 2>&1 | tail -1 | tee gpurun_out/${TAG}_soak.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
