#!/bin/bash
# dev helper (GPU box), round 2 run H: GPU tier, kernel-only timings of the lookup kernels after the 32-bit window minima, one
# ncu --set full capture of k_pseudoalign_small with a 160-line hot-line table
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
BIG=${BIG:-synth_4546_big}
if [ -z "$SKIP_TESTS" ]; then timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4; fi
ko() { # label, -- bench args
  local label=$1; shift; shift
  timeout 300 python bench.py --kernel-only --steps 5 --warmup 3 "$@" 2>>gpurun_out/ab.err | python -c "
import json,sys
j=json.loads(sys.stdin.read()); c=j['configs'][0]
print('$label', '%.1f M reads/s' % (c['value']/1e6), {k: round(v,3) for k,v in c['kernel_ms'].items()})" | tee -a gpurun_out/ab_h.txt
}
: > gpurun_out/ab_h.txt
ko s10_fi -- --reads 10000000
ko s10_tu -- --reads 10000000 --algo tu
ko big_fi -- --index $BIG.fur --reads 500000
ko big_mfur_tu_mixed -- --index $BIG.mfur --reads 500000 --algo tu --min-len 75 --max-len 300
NCU="ncu --clock-control none"
timeout 600 $NCU --set full --import-source on -k "regex:k_pseudoalign_small" -s 3 -c 1 -o gpurun_out/prof_s10_fi -f python bench.py --kernel-only --steps 1 --warmup 3 --reads 2000000 > gpurun_out/ncu_prof_s10_fi.log 2>&1
tail -1 gpurun_out/ncu_prof_s10_fi.log | cut -c1-160
ncu -i gpurun_out/prof_s10_fi.ncu-rep --page raw --csv > gpurun_out/prof_s10_fi_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_s10_fi.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/prof_s10_fi_src.csv 2>/dev/null
python tools/ncu_lines.py gpurun_out/prof_s10_fi_src.csv 160 > gpurun_out/prof_s10_fi_hot_lines.txt 2>&1
python tools/ncu_opmix.py gpurun_out/prof_s10_fi_src.csv > gpurun_out/prof_s10_fi_opmix.txt 2>&1
rm -f gpurun_out/*.ncu-rep gpurun_out/*_src.csv
du -sh gpurun_out
