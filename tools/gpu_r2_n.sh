#!/bin/bash
# dev helper (GPU box with N GPUs): the bench line at N ranks (torchrun, like the driver launches it) and, at N >= 2, the shipped tool's --gpus 2 test
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}; STEPS=${2:-10}
if [ "$N" -ge 2 ] && [ -z "$SKIP_TESTS" ]; then timeout 600 python -m pytest tests/test_cli.py -m gpu -x -q -k "two_real_gpus" 2>&1 | tail -3; fi
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps $STEPS --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "torchrun rc=$? json bytes=$(wc -c < gpurun_out/bench_n$N.json)"
grep -v "OMP_NUM_THREADS\|^\*\*\*" gpurun_out/bench_n$N.err | tail -25 | cut -c1-300
python tools/bench_summary.py gpurun_out/bench_n$N.json
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv,noheader | head -8
nproc; free -g | head -2
