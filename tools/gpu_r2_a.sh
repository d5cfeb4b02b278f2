#!/bin/bash
# dev helper (GPU box), round 2 run A: GPU tier incl. the multi-partition / skew fixtures, benches and ncu captures on the
# salmonella_4546-scale stand-in (dictionary and table larger than L2)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
BIG=${BIG:-synth_4546_big}
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -16
for algo in fi tu; do
  timeout 400 python bench.py --index $BIG.fur --reads 1000000 --steps 5 --algo $algo --cpu-sample 4000 > gpurun_out/bench_big_$algo.json 2>> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench_big_$algo.json
done
timeout 400 python bench.py --index $BIG.mfur --reads 1000000 --steps 5 --algo tu --min-len 75 --max-len 300 --cpu-sample 4000 > gpurun_out/bench_big_mfur_tu_mixed.json 2>> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench_big_mfur_tu_mixed.json
tail -5 gpurun_out/bench.err
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file gpurun_out/launches_big_fi.csv python bench.py --index $BIG.fur --steps 2 --warmup 1 --reads 200000 --no-cpu-baseline > gpurun_out/ncu_big.log 2>&1
cap() { # name regex skip algo index extra
  timeout 600 $NCU --set full --import-source on -k regex:$2 -s $3 -c 1 -o gpurun_out/$1 -f python bench.py --index $5 --algo $4 --steps 1 --warmup 1 --reads 200000 --no-cpu-baseline $6 > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log | cut -c1-160
}
cap prof_big_k1 k_fetch_color_sets 1 fi $BIG.fur
cap prof_big_k2fi k_color_sets_table 1 fi $BIG.fur
cap prof_big_k2tu k_color_sets_table 1 tu $BIG.fur
cap prof_big_emit k_emit_bits 1 fi $BIG.fur
cap prof_big_k2tu_mfur_mixed k_color_sets_table 1 tu $BIG.mfur "--min-len 75 --max-len 300"
FULGOR_GPU_TABLE_MAX_MB=0 cap prof_big_k2general k_color_sets_general 1 fi $BIG.fur
ls -la gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc
