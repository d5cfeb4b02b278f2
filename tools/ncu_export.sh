#!/bin/bash
# dev helper (GPU box): turn gpurun_out/*.ncu-rep into the small text artefacts that travel back (raw metrics page per kernel,
# per-source-line instruction / stall table) and delete the reports (gpurun_out/ is capped at 64 MiB)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for rep in gpurun_out/*.ncu-rep; do
  [ -f "$rep" ] || continue
  base=${rep%.ncu-rep}
  ncu -i "$rep" --page raw --csv > "${base}_raw.csv" 2>/dev/null
  ncu -i "$rep" --page source --csv --print-source cuda,sass > "${base}_src.csv" 2>/dev/null
  python tools/ncu_lines.py "${base}_src.csv" ${HOT_LINES:-45} > "${base}_hot_lines.txt" 2>&1
  python tools/ncu_opmix.py "${base}_src.csv" > "${base}_opmix.txt" 2>&1
  rm -f "$rep" "${base}_src.csv"
done
