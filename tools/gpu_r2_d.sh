#!/bin/bash
# dev helper (GPU box), round 2 run D: A/B runs of kernel variants (--kernel-only benches) + the parity tests the changes touch
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
BIG=${BIG:-synth_4546_big}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dedup or 4546_color_standin or walk or packed or without" 2>&1 | tail -5
ko() { # label, env..., -- bench args
  local label=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --kernel-only --steps 5 --warmup 3 "$@" 2>>gpurun_out/ab.err | python -c "
import json,sys
j=json.loads(sys.stdin.read()); c=j['configs'][0]
print('$label', '%.1f M reads/s' % (c['value']/1e6), {k: round(v,3) for k,v in c['kernel_ms'].items()})" | tee -a gpurun_out/ab.txt
}
: > gpurun_out/ab.txt
ko s10_fi_mb4 X=1 -- --reads 4000000
for mb in 3 5 6; do ko s10_fi_mb$mb FULGOR_GPU_LIB=build/libfulgor_gpu_mb$mb.so -- --reads 4000000; done
ko big_fi_ring X=1 -- --index $BIG.fur --reads 500000
ko big_fi_emit_v1 FULGOR_GPU_EMIT=1 -- --index $BIG.fur --reads 500000
ko big_fi_prefetch_mb5 FULGOR_GPU_LIB=build/libfulgor_gpu_mb5.so -- --index $BIG.fur --reads 500000
ko big_fi_prefetch_mb3 FULGOR_GPU_LIB=build/libfulgor_gpu_mb3.so -- --index $BIG.fur --reads 500000
ko big_tu_ring X=1 -- --index $BIG.fur --reads 500000 --algo tu
ko big_tu_prefetch FULGOR_GPU_LIB=build/libfulgor_gpu_mb5.so -- --index $BIG.fur --reads 500000 --algo tu
ko big_mfur_tu_mixed_ring X=1 -- --index $BIG.mfur --reads 500000 --algo tu --min-len 75 --max-len 300
ko big_mfur_tu_mixed_prefetch FULGOR_GPU_LIB=build/libfulgor_gpu_mb5.so -- --index $BIG.mfur --reads 500000 --algo tu --min-len 75 --max-len 300
tail -3 gpurun_out/ab.err
NCU="ncu --clock-control none"
timeout 600 $NCU --set full --import-source on -k "regex:k_color_sets_table" -s 3 -c 1 -o gpurun_out/prof_big_tu_ring -f python bench.py --kernel-only --steps 1 --warmup 3 --index $BIG.fur --reads 200000 --algo tu > gpurun_out/ncu_tu_ring.log 2>&1
bash tools/ncu_export.sh
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out
