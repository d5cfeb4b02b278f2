#!/bin/bash
# dev helper (GPU box): ncu full capture of k_fetch_color_sets and k_color_sets_table on the 4,546-color stand-in (big launch)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fetch_color_sets -s 3 -c 1 -o gpurun_out/prof_k1f -f python bench.py --index synth_4546.fur --steps 1 --warmup 1 --reads 1000000 --no-cpu-baseline > gpurun_out/ncu_k1f.log 2>&1
tail -2 gpurun_out/ncu_k1f.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_color_sets_table -s 3 -c 1 -o gpurun_out/prof_k2t -f python bench.py --index synth_4546.fur --steps 1 --warmup 1 --reads 1000000 --no-cpu-baseline > gpurun_out/ncu_k2t.log 2>&1
tail -2 gpurun_out/ncu_k2t.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
