#!/bin/bash
# fresh ncu --set full capture of the headline kernel (k_qsquat<float,3>) on the final sources: roofline.traffic stamp
set -u
TAG=${1:-r2au}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_qsquat -s 101 -c 1 -f -o /tmp/${TAG}_squat_osc \
  python bench.py --workload squat_osc --steps 2 --warmup 1 --preadvance 1000 --no-cpu-baseline > gpurun_out/${TAG}_squat_osc_ncu.log 2>&1
python tools/summarize_ncu.py /tmp/${TAG}_squat_osc.ncu-rep > gpurun_out/${TAG}_squat_osc.txt 2>&1
python tools/summarize_ncu.py /tmp/${TAG}_squat_osc.ncu-rep --traffic > gpurun_out/${TAG}_squat_osc_traffic.txt 2>&1
cat gpurun_out/${TAG}_squat_osc_traffic.txt
head -40 gpurun_out/${TAG}_squat_osc.txt
echo "== rollout kernel (OSC action space), one launch"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_qrollout -s 2 -c 1 -f -o /tmp/${TAG}_rollout_osc \
  python tools/bench_rollout.py --mode OSC --T 20 --reps 2 > gpurun_out/${TAG}_rollout_osc_ncu.log 2>&1
python tools/summarize_ncu.py /tmp/${TAG}_rollout_osc.ncu-rep > gpurun_out/${TAG}_rollout_osc.txt 2>&1
head -40 gpurun_out/${TAG}_rollout_osc.txt
