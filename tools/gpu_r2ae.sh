#!/bin/bash
set -u
TAG=${1:-r2ae}
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/${TAG}_smoke.txt
echo "== bench3d"; timeout 900 python tools/bench3d.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/${TAG}_bench3d.json; cut -c1-200 gpurun_out/${TAG}_bench3d.json
echo "== headline"; timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${TAG}_bench.json; cut -c1-200 gpurun_out/${TAG}_bench.json
