#!/bin/bash
# dev helper (GPU box): tool-level timing of the per-k-mer tools next to the reference's own (2 M reads, all host threads)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=$(nproc); W=/dev/shm/fg_tools_bench; mkdir -p $W gpurun_out build
[ -f $W/s10.gpk ] || xz -dkc data/salmonella_10.gpk.xz > $W/s10.gpk
[ -x build/readgen ] || g++ -O2 -std=c++17 -pthread -DREADGEN_MAIN tools/readgen.cpp -o build/readgen
build/readgen $W/s10.gpk ${1:-2000000} $W/reads.fq
OUT=gpurun_out/tools_bench.txt; : > $OUT
for tool in kmer-conservation kmer-matches; do
  timeout 60 ./fulgor_b200/fulgor_b200_pseudoalign $tool -i data/salmonella_10.fur -q $W/reads.fq -o $W/gpu.$tool -t $T --verbose | grep -E "elapsed" | sed "s/^/gpu $tool: /" | tee -a $OUT
  timeout 90 oracle/_ref/fulgor_ref $tool -i data/salmonella_10.fur -q $W/reads.fq -o $W/ref.$tool -t $T --verbose 2>&1 | grep -E "elapsed" | sed "s/^/ref $tool (-t $T): /" | tee -a $OUT
done
sort $W/gpu.kmer-conservation | md5sum | sed 's/^/gpu kmer-conservation sorted md5: /' | tee -a $OUT
sort $W/ref.kmer-conservation | md5sum | sed 's/^/ref kmer-conservation sorted md5: /' | tee -a $OUT
ls -la $W | tee -a $OUT
rm -rf $W
