#!/bin/bash
# dev helper (GPU box), round 2 run J: A/B of kernel variants (fulgor_b200/variants/*.so through FULGOR_GPU_LIB) on the kernel-only
# timings, the GPU parity tests of the color-set path, the default bench line with its wall time, region tables of K1
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
BIG=${BIG:-synth_4546_big}
ko() { # label, -- bench args
  local label=$1; shift; shift
  timeout 300 python bench.py --kernel-only --steps 5 --warmup 3 "$@" 2>>gpurun_out/ab.err | python -c "
import json,sys
j=json.loads(sys.stdin.read()); c=j['configs'][0]
print('$label', '%.1f M reads/s' % (c['value']/1e6), {k: round(v,3) for k,v in c['kernel_ms'].items()})" | tee -a gpurun_out/ab_j.txt
}
: > gpurun_out/ab_j.txt; : > gpurun_out/ab.err
for v in "" $VARIANTS; do
  if [ -n "$v" ]; then export FULGOR_GPU_LIB=$PWD/fulgor_b200/variants/libfulgor_gpu_$v.so; else unset FULGOR_GPU_LIB; fi
  tag=${v:-new}
  [ -n "$S10" ] && ko s10_fi_$tag -- --reads 10000000
  ko big_fi_$tag -- --index $BIG.fur --reads 500000
  [ -z "$FI_ONLY" ] && ko big_tu_$tag -- --index $BIG.fur --reads 500000 --algo tu
  [ -z "$FI_ONLY" ] && ko big_mfur_tu_mixed_$tag -- --index $BIG.mfur --reads 500000 --algo tu --min-len 75 --max-len 300
done
for v in $VARIANTS_FI; do
  export FULGOR_GPU_LIB=$PWD/fulgor_b200/variants/libfulgor_gpu_$v.so
  ko big_fi_$v -- --index $BIG.fur --reads 500000
done
for v in $VARIANTS_TU; do
  export FULGOR_GPU_LIB=$PWD/fulgor_b200/variants/libfulgor_gpu_$v.so
  ko big_tu_$v -- --index $BIG.fur --reads 500000 --algo tu
  ko big_mfur_tu_mixed_$v -- --index $BIG.mfur --reads 500000 --algo tu --min-len 75 --max-len 300
done
unset FULGOR_GPU_LIB
if [ "$TESTS" = all ]; then timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4; TESTS=""; fi
if [ -n "$TESTS" ]; then timeout 900 python -m pytest tests -m gpu -x -q -k "$TESTS" 2>&1 | tail -3; fi
if [ -n "$BENCH" ]; then
  t0=$(date +%s); timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? wall $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/bench.err
  python tools/bench_summary.py gpurun_out/bench.json
  t0=$(date +%s); timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; echo "ref rc=$? wall $(( $(date +%s) - t0 )) s"; cut -c1-300 gpurun_out/bench_ref.json
fi
NCU="ncu --clock-control none"
KO="--kernel-only --steps 1 --warmup 3"
cap() { # name regex skip count regions... -- bench-args...
  local name=$1 re=$2 skip=$3 cnt=$4 reads=$5; shift 5
  timeout 600 $NCU --set full --import-source on -k "regex:$re" -s $skip -c $cnt -o gpurun_out/$name -f python bench.py $KO --reads $reads "$@" > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log | cut -c1-160
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${name}_src.csv 2>/dev/null
  python tools/ncu_lines.py gpurun_out/${name}_src.csv ${HOT_LINES:-60} > gpurun_out/${name}_hot_lines.txt 2>&1
  python tools/ncu_opmix.py gpurun_out/${name}_src.csv > gpurun_out/${name}_opmix.txt 2>&1
  [ -n "$REGIONS" ] && python tools/ncu_regions.py gpurun_out/${name}_src.csv $reads $REGIONS > gpurun_out/${name}_regions.txt 2>&1
  rm -f gpurun_out/$name.ncu-rep gpurun_out/${name}_src.csv
}
for c in $CAPS; do
  case $c in
    big_fi_k1) cap prof_big_fi_k1 k_fetch_color_sets 3 1 200000 --index $BIG.fur ;;
    big_fi_k2) cap prof_big_fi_k2 k_color_sets_table 3 1 200000 --index $BIG.fur ;;
    big_fi_emit) cap prof_big_fi_emit k_emit_bits 3 1 200000 --index $BIG.fur ;;
    big_tu_k2) cap prof_big_tu_k2 k_color_sets_table 3 1 200000 --index $BIG.fur --algo tu ;;
    mfur_k2) cap prof_big_mfur_tu_mixed_k2 k_color_sets_table 3 1 200000 --index $BIG.mfur --algo tu --min-len 75 --max-len 300 ;;
    mfur_k1) cap prof_big_mfur_tu_mixed_k1 k_fetch_color_sets 3 1 200000 --index $BIG.mfur --algo tu --min-len 75 --max-len 300 ;;
    s10_fi) cap prof_s10_fi k_pseudoalign_small 3 1 2000000 ;;
  esac
done
du -sh gpurun_out
