#!/bin/bash
# dev helper (GPU box): end-of-round verification -- full GPU tier, smoke, benches (both arms, both algorithms, stand-ins),
# compute-sanitizer, one ncu capture of the threshold-union table kernel. Most important first: the call may be cut short.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=12 2>&1 | tail -22
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; cut -c1-200 gpurun_out/bench.json
timeout 300 python bench.py --algo tu > gpurun_out/bench_tu.json 2>> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench_tu.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench_ref.json
if [ -f fixtures_big/synth_4546.fur ]; then
  for algo in fi tu; do
    timeout 300 python bench.py --index synth_4546.fur --reads 1000000 --steps 5 --algo $algo --cpu-sample 4000 > gpurun_out/bench_big_$algo.json 2>> gpurun_out/bench.err; cut -c1-160 gpurun_out/bench_big_$algo.json
  done
  timeout 300 python bench.py --index synth_4546.mfur --reads 1000000 --steps 5 --algo tu --min-len 75 --max-len 300 --cpu-sample 4000 > gpurun_out/bench_big_mfur_tu_mixed.json 2>> gpurun_out/bench.err; cut -c1-160 gpurun_out/bench_big_mfur_tu_mixed.json
fi
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 300 $CS --tool $tool --print-limit 20 python tools/sanitize_smoke.py 150 > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitizer_$tool.log
done
if [ -f fixtures_big/synth_4546.fur ]; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_color_sets_table -s 3 -c 1 -o gpurun_out/prof_k2tu -f python bench.py --index synth_4546.fur --algo tu --steps 1 --warmup 1 --reads 1000000 --no-cpu-baseline > gpurun_out/ncu_k2tu.log 2>&1; tail -1 gpurun_out/ncu_k2tu.log | cut -c1-200
fi
