#!/bin/bash
# Runs ON THE GPU BOX: compute-sanitizer over the update kernels (SURVEY §4 / §5 "race detection"). Logs -> gpurun_out/${TAG}_san_*.log
TAG=${1:-r2}
shift
WHAT=${@:-row wide}
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py $WHAT > gpurun_out/${TAG}_san_${tool}.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target done" gpurun_out/${TAG}_san_${tool}.log | tail -3
done
