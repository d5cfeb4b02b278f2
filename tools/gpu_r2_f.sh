#!/bin/bash
# dev helper (GPU box), round 2 run F: the multi-config bench line (both arms), launch lists, ncu --set full captures of every
# kernel the bench names (from --kernel-only runs: fixed launch sequence), compute-sanitizer smoke. Reports are exported to text
# and deleted (gpurun_out/ is capped at 64 MiB).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
BIG=${BIG:-synth_4546_big}
if [ -n "$RUN_TESTS" ]; then timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4; fi
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench_ref.json
NCU="ncu --clock-control none"
KO="--kernel-only --steps 1 --warmup 3"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file gpurun_out/launches_s10.csv python bench.py --only-primary --steps 2 --warmup 3 --reads 2000000 --no-cpu-baseline > gpurun_out/ncu_l1.log 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file gpurun_out/launches_big.csv python bench.py --index $BIG.fur --steps 2 --warmup 3 --reads 200000 --no-cpu-baseline > gpurun_out/ncu_l2.log 2>&1
cap() { # name regex skip count bench-args...
  local name=$1 re=$2 skip=$3 cnt=$4; shift 4
  timeout 600 $NCU --set full --import-source on -k "regex:$re" -s $skip -c $cnt -o gpurun_out/$name -f python bench.py $KO "$@" > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log | cut -c1-160
  bash tools/ncu_export.sh
}
cap prof_s10_fi k_pseudoalign_small 3 1 --reads 2000000
cap prof_big_fi_k1 k_fetch_color_sets 3 1 --index $BIG.fur --reads 200000
cap prof_big_fi_k2 k_color_sets_table 3 1 --index $BIG.fur --reads 200000
cap prof_big_fi_emit k_emit_bits 3 1 --index $BIG.fur --reads 200000
cap prof_big_tu_k2 k_color_sets_table 3 1 --index $BIG.fur --reads 200000 --algo tu
cap prof_big_mfur_tu_mixed_k2 k_color_sets_table 3 1 --index $BIG.mfur --reads 200000 --algo tu --min-len 75 --max-len 300
FULGOR_GPU_TABLE_MAX_MB=0 cap prof_big_general k_color_sets_general 3 1 --index $BIG.fur --reads 100000
rm -f gpurun_out/*.ncu-rep
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 600 $CS --tool $tool --print-limit 20 python tools/sanitize_smoke.py 150 > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitizer_$tool.log
done
du -sh gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
nproc; lscpu | grep "Model name"
