#!/bin/bash
# GPU call 10: compute-sanitizer (memcheck, racecheck, synccheck) over every solve kernel on tiny fits.
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== $tool"; timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_RUN_OK|hazard|Invalid|error" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -12
done
