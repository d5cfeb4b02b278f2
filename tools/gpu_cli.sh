#!/bin/bash
# dev helper (GPU box): CLI-level timing on salmonella_10, 10 M reads, 1 GPU, all host threads
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nproc
bash tools/cli_bench.sh ${1:-10000000} data/salmonella_10.fur salmonella_10 1 2>&1 | tail -14
cp gpurun_out/cli_bench.txt gpurun_out/cli_bench_1gpu.txt
