#!/bin/bash
# compute-sanitizer on the small configs (scripts/sanitize.py): memcheck, racecheck, synccheck, initcheck summaries
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize run ok|Error|hazard" gpurun_out/sanitizer_$tool.log | head -8
done
