#!/bin/bash
set -u
TAG=${1:-r2ad}
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool (3-D tree kernel)"
  timeout 1200 compute-sanitizer --tool $tool python tools/sanitize3d.py 2>&1 | grep -v "^$" | tail -25 | tee gpurun_out/${TAG}_sanitizer_${tool}.txt
done
