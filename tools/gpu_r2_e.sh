#!/bin/bash
# dev helper (GPU box), round 2 run E: parity tests + A/B of the lookup kernel's per-read table (hash vs append/sort) and of the
# table kernel's counters (pending vectors in shared memory, single pass) against the previous builds
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
BIG=${BIG:-synth_4546_big}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
ko() { # label, env..., -- bench args
  local label=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --kernel-only --steps 5 --warmup 3 "$@" 2>>gpurun_out/ab.err | python -c "
import json,sys
j=json.loads(sys.stdin.read()); c=j['configs'][0]
print('$label', '%.1f M reads/s' % (c['value']/1e6), {k: round(v,3) for k,v in c['kernel_ms'].items()})" | tee -a gpurun_out/ab2.txt
}
: > gpurun_out/ab2.txt
for v in new head oldk1; do
  LIBENV="X=1"; [ $v != new ] && LIBENV="FULGOR_GPU_LIB=build/libfulgor_gpu_$v.so"
  ko big_fi_$v $LIBENV -- --index $BIG.fur --reads 500000
  ko big_tu_$v $LIBENV -- --index $BIG.fur --reads 500000 --algo tu
  ko big_mfur_tu_mixed_$v $LIBENV -- --index $BIG.mfur --reads 500000 --algo tu --min-len 75 --max-len 300
done
ko s10_fi_new X=1 -- --reads 4000000
tail -3 gpurun_out/ab.err
du -sh gpurun_out
