#!/bin/bash
# dev helper (GPU box): the per-read list change -- parity on the many-color indexes + stand-in benches
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "4546 or long_reads or fetch or dedup or multi_chunk or kmer" 2>&1 | tail -4
for algo in fi tu; do
  timeout 600 python bench.py --index synth_4546.fur --reads 1000000 --steps 5 --algo $algo --cpu-sample 4000 > gpurun_out/bench_big_$algo.json 2>> gpurun_out/bench_big.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_big_$algo.json')); r=d['roofline']; print('$algo', round(d['value']/1e6,1), 'M reads/s; lookup', round(r['lookup_ms'],2), 'color sets', round(r['color_sets_ms'],2), 'emit', round(r['scan_emit_ms'],2), 'frac', round(r['frac'],3), 'e2e', round(d['e2e']['value']/1e6,2), d['cpu_baseline']['matches_gpu_output'])"
done
timeout 600 python bench.py --index synth_4546.mfur --reads 1000000 --steps 5 --algo tu --min-len 75 --max-len 300 --cpu-sample 4000 > gpurun_out/bench_big_mfur_tu_mixed.json 2>> gpurun_out/bench_big.err
python -c "
import json; d=json.load(open('gpurun_out/bench_big_mfur_tu_mixed.json')); r=d['roofline']; print('mfur tu mixed', round(d['value']/1e6,1), 'M reads/s; lookup', round(r['lookup_ms'],2), 'color sets', round(r['color_sets_ms'],2), 'emit', round(r['scan_emit_ms'],2), d['cpu_baseline']['matches_gpu_output'])"
tail -3 gpurun_out/bench_big.err
