#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_driver.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|SANITIZE_DRIVER|RACECHECK SUMMARY" gpurun_out/sanitizer_$tool.log | tail -3
done
