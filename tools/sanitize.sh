#!/bin/bash
# compute-sanitizer runs of tools/sanitize_target.py; summaries to gpurun_out/sanitizer_*.txt
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?" > gpurun_out/sanitizer_$tool.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok|Error|hazard" gpurun_out/sanitizer_$tool.log | sort | uniq -c | head -30 >> gpurun_out/sanitizer_$tool.txt
done
cat gpurun_out/sanitizer_*.txt
