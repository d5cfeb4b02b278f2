#!/bin/bash
# dev helper (GPU box): parity + bench + ncu launch list on a 4,546-color stand-in (needs fixtures_big/ in the snapshot)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
IDX=${1:-synth_4546_dense.fur}; READS=${2:-200000}
timeout 900 python -m pytest tests/test_cli.py tests/test_gpu_parity.py -m gpu -x -q -k "cli or 4546" 2>&1 | tail -5
timeout 900 python bench.py --index $IDX --reads $READS --steps 5 --cpu-sample 4000 > gpurun_out/bench_big.json 2> gpurun_out/bench_big.err; tail -3 gpurun_out/bench_big.err; cat gpurun_out/bench_big.json
timeout 900 python bench.py --index $IDX --reads $READS --steps 5 --algo tu --no-cpu-baseline > gpurun_out/bench_big_tu.json 2>> gpurun_out/bench_big.err; cat gpurun_out/bench_big_tu.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_big.csv python bench.py --index $IDX --steps 2 --warmup 1 --reads 100000 --no-cpu-baseline > gpurun_out/ncu_big.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_color_sets_general -s 1 -c 1 -o gpurun_out/prof_k2 -f python bench.py --index $IDX --steps 1 --warmup 1 --reads 100000 --no-cpu-baseline > gpurun_out/ncu_full_big.log 2>&1
grep -c k_ gpurun_out/launches_big.csv
