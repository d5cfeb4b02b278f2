#!/bin/bash
# dev helper (GPU box with N GPUs): the driver's scaling run for one N, both arms
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-8}; mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv,noheader | head -8; nproc; free -g | head -2
nvidia-smi topo -m 2>&1 | head -14 | cut -c1-220
lscpu | grep -i "numa\|socket\|model name"
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && [ "$(cat $d/class 2>/dev/null | cut -c1-6)" = "0x0302" ]; then echo "$(basename $d) numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist)"; fi; done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -5 gpurun_out/bench_n$N.err; cut -c1-1200 gpurun_out/bench_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
