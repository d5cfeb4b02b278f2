#!/bin/bash
# dev helper (GPU box): second half of the end-of-round verification (after the CLI fix and the carry-save restructuring)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_cli.py -m gpu -q -x -k kmer_tools 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_oracle_golden.py -m gpu -q -x -k "not 4546" --durations=8 2>&1 | tail -14
timeout 200 python bench.py --index synth_4546.fur --reads 1000000 --steps 5 --algo tu --cpu-sample 4000 > gpurun_out/bench_big_tu.json 2>> gpurun_out/bench.err; cut -c1-160 gpurun_out/bench_big_tu.json
timeout 200 python bench.py --index synth_4546.mfur --reads 1000000 --steps 5 --algo tu --min-len 75 --max-len 300 --cpu-sample 4000 > gpurun_out/bench_big_mfur_tu_mixed.json 2>> gpurun_out/bench.err; cut -c1-160 gpurun_out/bench_big_mfur_tu_mixed.json
timeout 200 python bench.py --index synth_4546.mdfur --reads 500000 --steps 3 --algo tu --threshold 0.5 --cpu-sample 3000 > gpurun_out/bench_big_mdfur_tu.json 2>> gpurun_out/bench.err; cut -c1-160 gpurun_out/bench_big_mdfur_tu.json
tail -2 gpurun_out/bench.err
