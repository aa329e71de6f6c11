#!/bin/bash
# compute-sanitizer memcheck + racecheck + synccheck over tools/gpu/sanitize_workload.py; summaries -> gpurun_out/sanitizer_*.log
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/gpu/sanitize_workload.py > gpurun_out/sanitizer_$tool.full.log 2>&1
  echo "== $tool: exit $?" | tee gpurun_out/sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload|Error|hazard" gpurun_out/sanitizer_$tool.full.log | sort | uniq -c | sort -rn | head -12 | tee -a gpurun_out/sanitizer_$tool.log
done
