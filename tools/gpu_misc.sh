#!/bin/bash
# dev helper (GPU box): PCIe probe, deduplication timing, ncu launch list of the many-color path with dedup / kmer tools
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python tools/pcie_probe.py > gpurun_out/bw_test.txt 2>&1; grep -v "^$" gpurun_out/bw_test.txt | head -14
timeout 600 python tools/dedup_bench.py salmonella_10.fur 4000000 200000 5 > gpurun_out/dedup_s10.json 2> gpurun_out/dedup.err; cat gpurun_out/dedup_s10.json
timeout 600 python tools/dedup_bench.py synth_4546.fur 1000000 50000 5 > gpurun_out/dedup_synth4546.json 2>> gpurun_out/dedup.err; cat gpurun_out/dedup_synth4546.json
timeout 600 python tools/dedup_bench.py synth_4546.mdfur 1000000 50000 5 > gpurun_out/dedup_synth4546_mdfur.json 2>> gpurun_out/dedup.err; cat gpurun_out/dedup_synth4546_mdfur.json
tail -3 gpurun_out/dedup.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_dedup.csv python tools/dedup_bench.py synth_4546.fur 200000 10000 1 > /dev/null 2>&1
grep -c . gpurun_out/launches_dedup.csv
