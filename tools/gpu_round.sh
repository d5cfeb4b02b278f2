#!/bin/bash
# dev helper, run on the GPU box via gpurun: tests + smoke + bench (both arms) + ncu launch list + ncu full capture of the top kernel
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest ${GPU_TESTS:-tests} -m gpu -x -q 2>&1 | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
timeout 900 python bench.py --algo tu > gpurun_out/bench_tu.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_tu.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --reads 2000000 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pseudoalign_small -s 1 -c 1 -o gpurun_out/prof_k1 -f python bench.py --steps 1 --warmup 1 --reads 2000000 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
python tools/pcie_probe.py 2>&1 | tail -4
ls -la gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc; lscpu | grep "Model name"
