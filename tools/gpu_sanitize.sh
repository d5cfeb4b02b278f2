#!/bin/bash
# dev helper (GPU box): compute-sanitizer (memcheck, racecheck, synccheck, initcheck) over a small pass of every kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 $CS --tool $tool --print-limit 20 python tools/sanitize_smoke.py 200 > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|Error|hazard" gpurun_out/sanitizer_$tool.log | sort | uniq -c | sort -rn | head -12
done
