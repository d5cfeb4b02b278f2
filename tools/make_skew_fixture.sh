#!/bin/bash
# Fixture tooling (not on the query path): builds data/synth_skew.fur (+ data/synth_skew.gpk.xz), the small COMMITTED index that
# takes the dictionary branches salmonella_10 / synth_200 never reach:
#   - a minimizer MPHF with SEVERAL partitions (pthash partitioned_phf.hpp:155-159). SSHash partitions at 3,000,000 minimizers
#     (external/sshash/include/constants.hpp:13), far beyond a committable file, so the BUILDER used here is the reference
#     compiled from a scratch copy in which that one constant is 20,000. The file format carries the partition count, so the
#     UNMODIFIED reference (oracle/_ref) loads and queries the result like any other index -- which is what the tests pin against;
#   - a skew index with all seven size classes (skew_index.hpp:40-52), one of them EMPTY, the last one absorbing buckets beyond
#     2^max_l and itself a partitioned MPHF: tools/synthgen.py --plant embeds minimizer-winning 20-mers in thousands of contexts;
#   - canonical minimizers that are their own reverse complement and are read off one strand only (--plant-palindromes, possible
#     for even m): the orientation of the indexed string cannot be inferred from how the minimizer reads.
# Needs /root/reference. About two minutes (most of it compiling the builder).
set -euo pipefail
cd "$(dirname "$0")/.."
REF=${REF:-/root/reference}; TMP=${TMPDIR:-/tmp}/fg_skew_fixture
mkdir -p "$TMP/ref" "$TMP/out" build
cp -r "$REF/." "$TMP/ref/"
sed -i 's/avg_partition_size = 3000000/avg_partition_size = 20000/' "$TMP/ref/external/sshash/include/constants.hpp"
make -C oracle "$TMP/out/fulgor_ref" OUT="$TMP/out" REF="$TMP/ref"
[ -x build/mkdump ] && [ build/mkdump -nt tools/mkdump.cpp ] || g++ -O2 -std=c++17 tools/mkdump.cpp -o build/mkdump -lz
python tools/synthgen.py "$TMP/genomes" 16 100000 --seed 11 --sub 0.002 --hgt 0.5 --novel 30000 40000 --plant 20 --plant-palindromes --plant-targets 80,160,320,640,1300,2600,7500
build/mkdump "$TMP/synth_skew" "@$TMP/genomes/list.txt"
"$TMP/out/fulgor_ref" load -i "$TMP/synth_skew" -o "$TMP/synth_skew" -m 20 -d "$TMP" -t 8 --verbose | grep -i "partitions\|partition_id"
cp "$TMP/synth_skew.fur" data/synth_skew.fur
xz -9 -c "$TMP/synth_skew.gpk" > data/synth_skew.gpk.xz
ls -la data/synth_skew.*
