#!/bin/bash
# dev helper (GPU box): full GPU test tier + stand-in benches incl. differential types and --deduplicate
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for idx in synth_4546.dfur synth_4546.mdfur; do
  timeout 600 python bench.py --index $idx --reads 1000000 --steps 5 --algo fi --cpu-sample 4000 > gpurun_out/bench_big_${idx#*.}_fi.json 2>> gpurun_out/bench_big.err; cut -c1-1500 gpurun_out/bench_big_${idx#*.}_fi.json
done
tail -5 gpurun_out/bench_big.err
