#!/bin/bash
# dev helper (GPU box), round 2 run I: the default bench line at the round's last kernels, kernel-only timings on the big
# stand-in, ncu --set full captures of the lookup kernel (k_fetch_color_sets) on the big stand-in (hybrid FI, meta TU mixed)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
BIG=${BIG:-synth_4546_big}
/usr/bin/time -f "bench wall %e s" timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json
ko() { # label, -- bench args
  local label=$1; shift; shift
  timeout 300 python bench.py --kernel-only --steps 5 --warmup 3 "$@" 2>>gpurun_out/ab.err | python -c "
import json,sys
j=json.loads(sys.stdin.read()); c=j['configs'][0]
print('$label', '%.1f M reads/s' % (c['value']/1e6), {k: round(v,3) for k,v in c['kernel_ms'].items()})" | tee -a gpurun_out/ab_i.txt
}
: > gpurun_out/ab_i.txt
ko big_fi -- --index $BIG.fur --reads 500000
ko big_tu -- --index $BIG.fur --reads 500000 --algo tu
ko big_mfur_tu_mixed -- --index $BIG.mfur --reads 500000 --algo tu --min-len 75 --max-len 300
NCU="ncu --clock-control none"
KO="--kernel-only --steps 1 --warmup 3"
cap() { # name regex skip count bench-args...
  local name=$1 re=$2 skip=$3 cnt=$4; shift 4
  timeout 600 $NCU --set full --import-source on -k "regex:$re" -s $skip -c $cnt -o gpurun_out/$name -f python bench.py $KO "$@" > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log | cut -c1-160
  bash tools/ncu_export.sh
}
cap prof_big_fi_k1 k_fetch_color_sets 3 1 --index $BIG.fur --reads 200000
cap prof_big_mfur_tu_mixed_k1 k_fetch_color_sets 3 1 --index $BIG.mfur --reads 200000 --algo tu --min-len 75 --max-len 300
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out
