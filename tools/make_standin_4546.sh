#!/bin/bash
# Fixture tooling (not on the query path): builds the labelled STAND-IN for salmonella_4546 (the real 4,546 genomes are a Zenodo
# download, unavailable offline): 4,546 synthetic genomes evolved along a random tree with substitutions,
# indels and horizontal transfers (tools/synthgen.py), dumped to unitigs + color sets (tools/mkdump.cpp) and turned into a
# genuine .fur / .mfur / .dfur / .mdfur by the REFERENCE's own tools (oracle/_ref/fulgor_ref load -m 20, color --meta / --diff / --meta --diff), like README.md:158-160.
# Needs oracle/_ref (i.e. /root/reference at build time). Output: fixtures_big/NAME.{fur,mfur,dfur,mdfur,gpk} (git-ignored).
#   tools/make_standin_4546.sh [GENOME_LEN=200000] [N=4546] [SUB=0.00003] [NAME=synth_N]
# SYNTH_EXTRA (environment) is passed on to tools/synthgen.py. The salmonella_4546-SCALE stand-in (as many k-mers as the real
# index: ~46 M, several minimizer-MPHF partitions, a skew index with every size class, dictionary and decoded color-set table
# both larger than the 126 MB L2) is
#   SYNTH_EXTRA="--novel 4000 6000 --plant 20 --plant-scale 1.6" tools/make_standin_4546.sh 200000 4546 0.00003 synth_4546_big
# (about an hour on 8 cores, ~25 GB of RAM in mkdump).
# The default substitution rate gives ~0.4 unitigs per genome position, like the real collection (1.88 M unitigs over ~5 Mbp
# genomes, reference README.md:158-160); SUB=0.0005 with GENOME_LEN=100000 gives a much more fragmented stress case
# (fixtures_big/synth_4546_dense: ~7 unitigs per position, ~74 distinct color sets per 150 bp read).
set -euo pipefail
cd "$(dirname "$0")/.."
LEN=${1:-200000}; N=${2:-4546}; SUB=${3:-0.00003}; NAME=${4:-synth_$N}
OUT=fixtures_big; TMP=${TMPDIR:-/tmp}/fg_standin_$NAME
mkdir -p "$OUT" "$TMP" build
[ -x build/mkdump ] && [ build/mkdump -nt tools/mkdump.cpp ] || g++ -O2 -std=c++17 tools/mkdump.cpp -o build/mkdump -lz
python tools/synthgen.py "$TMP/genomes" "$N" "$LEN" --seed 4546 --sub "$SUB" --indel 0.10 --hgt 0.5 ${SYNTH_EXTRA:-}
build/mkdump "$TMP/$NAME" "@$TMP/genomes/list.txt"
oracle/_ref/fulgor_ref load -i "$TMP/$NAME" -o "$TMP/$NAME" -m 20 -d "$TMP" -t 8 --verbose
oracle/_ref/fulgor_ref color -i "$TMP/$NAME.fur" -d "$TMP" -t 8 --meta --verbose
# the differential and meta-differential re-encodings (about a minute each); their dictionaries are rebuilt (permuted unitigs)
oracle/_ref/fulgor_ref color -i "$TMP/$NAME.fur" -d "$TMP" -t 8 --diff --verbose
oracle/_ref/fulgor_ref color -i "$TMP/$NAME.fur" -d "$TMP" -t 8 --meta --diff --verbose
mkdir -p ${FG_FULL_DIR:-/tmp/fg_fixtures/full}
mv "$TMP/$NAME.fur" "$TMP/$NAME.mfur" "$TMP/$NAME.dfur" "$TMP/$NAME.mdfur" "$OUT/"
mv "$TMP/$NAME.gpk" ${FG_FULL_DIR:-/tmp/fg_fixtures/full}/
# reads are drawn from every 20th genome: a small file that travels to the GPU box (the full pack stays in ${FG_FULL_DIR:-/tmp/fg_fixtures/full}/)
python tools/gpk_subset.py "${FG_FULL_DIR:-/tmp/fg_fixtures/full}/$NAME.gpk" "$OUT/$NAME.gpk" 20
ls -la "$OUT"
