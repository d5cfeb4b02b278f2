#!/bin/bash
# dev helper (GPU box), last run of round 2: A/B of the lookup kernel (variants/prev = the commit before), then tools/gpu_r2_final.sh
# without sanitizer on the current library, captures of the lookup kernels only
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
S10=1 FI_ONLY=1 VARIANTS='prev' bash tools/gpu_r2_j.sh 2>&1 | grep "M reads/s"
BIG=${BIG:-synth_4546_big}
export ONLY_K1_CAPS=1 SKIP_SANITIZER=1
bash tools/gpu_r2_final.sh
