#!/bin/bash
# dev helper (GPU box): full GPU test tier
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --durations=20 2>&1 | tail -40
