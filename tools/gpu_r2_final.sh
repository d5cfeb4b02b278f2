#!/bin/bash
# dev helper (GPU box), end of round 2: ncu --set full captures of every kernel of the path (profiles/kernels.json is refreshed from
# them BEFORE the bench reads it), the whole GPU tier, smoke, the default bench line and the reference arm with their wall times,
# the launch list of a short bench run, compute-sanitizer. Everything lands in gpurun_out/ (copied to profiles/r02_* afterwards).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
BIG=${BIG:-synth_4546_big}
NCU="ncu --clock-control none"
KO="--kernel-only --steps 1 --warmup 3"
cap() { # name regex key reads bench-args...
  local name=$1 re=$2 key=$3 reads=$4; shift 4
  timeout 600 $NCU --set full --import-source on -k "regex:$re" -s 3 -c 1 -o gpurun_out/$name -f python bench.py $KO --reads $reads "$@" > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log | cut -c1-160
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${name}_src.csv 2>/dev/null
  python tools/ncu_lines.py gpurun_out/${name}_src.csv ${HOT_LINES:-45} > gpurun_out/${name}_hot_lines.txt 2>&1
  python tools/ncu_opmix.py gpurun_out/${name}_src.csv > gpurun_out/${name}_opmix.txt 2>&1
  python tools/ncu_kernels_json.py "$key" $reads gpurun_out/${name}_raw.csv profiles/r02_${name#prof_}_raw.csv | cut -c1-200
  rm -f gpurun_out/$name.ncu-rep gpurun_out/${name}_src.csv
}
if [ -z "$SKIP_CAPS" ]; then
cap prof_s10_fi k_pseudoalign_small "k_pseudoalign_small@salmonella_10.fur" 2000000
cap prof_big_fi_k1 k_fetch_color_sets "k_fetch_color_sets@$BIG.fur" 200000 --index $BIG.fur
if [ -z "$ONLY_K1_CAPS" ]; then
cap prof_big_fi_k2 k_color_sets_table "k_color_sets_table@$BIG.fur" 200000 --index $BIG.fur
cap prof_big_fi_emit k_emit_bits "k_emit_bits@$BIG.fur" 200000 --index $BIG.fur
cap prof_big_tu_k2 k_color_sets_table "k_color_sets_table[tu]@$BIG.fur" 200000 --index $BIG.fur --algo tu
fi
cap prof_big_mfur_tu_mixed_k1 k_fetch_color_sets "k_fetch_color_sets@$BIG.mfur" 200000 --index $BIG.mfur --algo tu --min-len 75 --max-len 300
[ -z "$ONLY_K1_CAPS" ] && cap prof_big_mfur_tu_mixed_k2 k_color_sets_table "k_color_sets_table@$BIG.mfur" 200000 --index $BIG.mfur --algo tu --min-len 75 --max-len 300
cp profiles/kernels.json gpurun_out/kernels.json
fi
if [ -z "$SKIP_TESTS" ]; then timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4; fi
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
t0=$(date +%s); timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? wall $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/bench.err
python tools/bench_summary.py gpurun_out/bench.json
t0=$(date +%s); timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; echo "ref rc=$? wall $(( $(date +%s) - t0 )) s"; cut -c1-200 gpurun_out/bench_ref.json
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_s10.csv python bench.py --steps 2 --warmup 1 --reads 2000000 --no-cpu-baseline --only-primary > gpurun_out/launches_bench.log 2>&1; echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches_s10.csv)"
if [ -z "$SKIP_SANITIZER" ]; then
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 400 $CS --tool $tool --print-limit 20 python tools/sanitize_smoke.py 150 > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitizer_$tool.log
done
fi
du -sh gpurun_out
