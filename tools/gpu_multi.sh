#!/bin/bash
# dev helper (GPU box with N GPUs): the bench under torchrun, N ranks, NCCL broadcast of the image
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}; mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -5 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 | tail -1 | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2>> gpurun_out/bench_n$N.err; cat gpurun_out/bench_n1.json | cut -c1-400
timeout 600 ./fulgor_b200/fulgor_b200_pseudoalign --help 2>&1 | head -2
bash tools/cli_bench.sh 10000000 data/salmonella_10.fur salmonella_10 $N 2>&1 | tail -12
