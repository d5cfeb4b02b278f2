#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
if [ "${1:-}" = "tests" ]; then timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "4546 or long_reads or fetch or dedup" 2>&1 | tail -3; fi
for idx in synth_4546.fur synth_200.fur; do
timeout 600 python bench.py --index $idx --reads 1000000 --steps 5 --algo fi --no-cpu-baseline > gpurun_out/b.json 2>> gpurun_out/bench_big.err
python -c "
import json; d=json.load(open('gpurun_out/b.json')); r=d['roofline']; print('$idx fi', round(d['value']/1e6,1), 'M reads/s; lookup', round(r['lookup_ms'],2), 'color sets', round(r['color_sets_ms'],2), 'emit', round(r['scan_emit_ms'],2), 'frac', round(r['frac'],3))"
done
tail -2 gpurun_out/bench_big.err
