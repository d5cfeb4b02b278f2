#!/bin/bash
# dev helper (GPU box): parity + bench + ncu on a 4,546-color stand-in (needs fixtures_big/ in the snapshot)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
IDX=${1:-synth_4546.fur}; READS=${2:-1000000}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "4546 or without" 2>&1 | tail -5
for algo in fi tu; do
  timeout 900 python bench.py --index $IDX --reads $READS --steps 5 --algo $algo --cpu-sample 4000 > gpurun_out/bench_big_$algo.json 2> gpurun_out/bench_big.err; tail -3 gpurun_out/bench_big.err; cat gpurun_out/bench_big_$algo.json
done
timeout 900 python bench.py --index ${IDX%.fur}.mfur --reads $READS --steps 5 --algo tu --min-len 75 --max-len 300 --cpu-sample 4000 > gpurun_out/bench_big_mfur_tu_mixed.json 2>> gpurun_out/bench_big.err; cat gpurun_out/bench_big_mfur_tu_mixed.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_big.csv python bench.py --index $IDX --steps 2 --warmup 1 --reads 200000 --no-cpu-baseline > gpurun_out/ncu_big.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_color_sets_table -s 1 -c 1 -o gpurun_out/prof_k2 -f python bench.py --index $IDX --steps 1 --warmup 1 --reads 200000 --no-cpu-baseline > gpurun_out/ncu_full_big.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fetch_color_sets -s 1 -c 1 -o gpurun_out/prof_k1f -f python bench.py --index $IDX --steps 1 --warmup 1 --reads 200000 --no-cpu-baseline > gpurun_out/ncu_full_big2.log 2>&1
