#!/bin/bash
# dev helper (GPU box): GPU tests, then kernel timing of library variants build/lib_*.so, then an ncu capture of the default build
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest ${GPU_AB_TESTS:-tests} -m gpu -x -q 2>&1 | tail -5
for lib in build/lib_*.so; do
  echo "== $lib"; FULGOR_GPU_LIB=$PWD/$lib timeout 300 python tools/quick_gpu.py 2000000 2>&1 | grep -v "^gen" | tail -9
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pseudoalign_small -s 1 -c 1 -o gpurun_out/prof_k1 -f python bench.py --steps 1 --warmup 1 --reads 2000000 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
