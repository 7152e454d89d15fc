#!/bin/bash
# phase-barrier A/B on the planar quad kernels (CASSIE_QUAD_PHASE_SYNC = 0 / 1 / 2) + the new 3-D NaN-guard test
set -u
TAG=${1:-r2z}
mkdir -p gpurun_out
echo "== pytest tree"; timeout 1200 python -m pytest tests/test_gpu_tree.py -m gpu -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest_tree.txt
bash tools/gpu_ab.sh ${TAG} "base ps1 ps2 base ps1 ps2" "squat_osc squat_jacobian" | tee gpurun_out/${TAG}_phase_sync_ab.txt
