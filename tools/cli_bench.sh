#!/bin/bash
# SURVEY.md 8(d) CLI-level timing on the GPU box: uncompressed FASTQ in /dev/shm, output to /dev/shm, the reference's own
# `elapsed` line for both tools (excludes index load), outputs compared after sorting by read id.
#   tools/cli_bench.sh [N_READS=10000000] [INDEX=data/salmonella_10.fur] [GPK=salmonella_10] [GPUS=1] [EXTRA_ARGS for both tools...]
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-10000000}; IDX=${2:-data/salmonella_10.fur}; GPK=${3:-salmonella_10}; GPUS=${4:-1}; if [ $# -ge 4 ]; then shift 4; else set --; fi; EXTRA="$*"
T=$(nproc); W=/dev/shm/fg_cli_bench; mkdir -p $W gpurun_out build
[ -f data/$GPK.gpk ] || [ -f fixtures_big/$GPK.gpk ] || xz -dkc data/$GPK.gpk.xz > $W/$GPK.gpk
GP=$( [ -f data/$GPK.gpk ] && echo data/$GPK.gpk || ( [ -f fixtures_big/$GPK.gpk ] && echo fixtures_big/$GPK.gpk || echo $W/$GPK.gpk ) )
[ -x build/readgen ] || g++ -O2 -std=c++17 -pthread -DREADGEN_MAIN tools/readgen.cpp -o build/readgen
build/readgen $GP $N $W/reads.fq
ls -la $W/reads.fq
OUT=gpurun_out/cli_bench.txt; : > $OUT
for rep in 1 2 3; do
  ./fulgor_b200/fulgor_b200_pseudoalign -i $IDX -q $W/reads.fq -o $W/gpu.out -t $T --gpus $GPUS --verbose $EXTRA | grep -E "elapsed|num_mapped" | sed "s/^/gpu  rep$rep (--gpus $GPUS -t $T): /" | tee -a $OUT
done
for rep in 1 2; do
  oracle/_ref/fulgor_ref pseudoalign -i $IDX -q $W/reads.fq -o $W/ref.out -t $T --verbose $EXTRA 2>&1 | grep -E "elapsed|num_mapped" | sed "s/^/ref  rep$rep (-t $T): /" | tee -a $OUT
done
sort -n $W/ref.out > $W/ref.sorted; if cmp -s $W/ref.sorted $W/gpu.out; then echo "outputs identical after sort by read id ($(wc -l < $W/gpu.out) records)" | tee -a $OUT; else echo "OUTPUTS DIFFER" | tee -a $OUT; fi
rm -rf $W
