#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python tools/quick_e2e.py 10000000 2097152 1048576 524288 262144 131072 2>&1 | tail -6
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "multi_chunk or E2BIG or e2big or dedup or empty" 2>&1 | tail -3
