"""dev helper: SASS opcode mix (warp instructions executed) from `ncu --page source --csv --print-source cuda,sass`.
usage: ncu_opmix.py src.csv n_reads"""
import csv, sys, collections
rows = csv.reader(open(sys.argv[1]))
n_reads = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
agg = collections.Counter(); seen = set(); hdr = None
for r in rows:
    if not r: continue
    if r[0] == 'Line No': hdr = r; ia = 2; isrc = 3; ii = r.index('Instructions Executed'); continue
    if hdr is None or len(r) <= ii: continue
    if r[0] == '' and r[ia].startswith('0x') and r[ia] not in seen:
        seen.add(r[ia])
        toks = r[isrc].split()
        if not toks: continue
        op = toks[1] if toks[0].startswith('@') else toks[0]
        op = op.split('.')[0]
        try: agg[op] += int(r[ii])
        except ValueError: pass
tot = sum(agg.values())
print("total warp inst", tot, "per read", tot / n_reads)
for k, v in agg.most_common(45): print(f"{k:12s} {v/tot*100:5.1f}%  {v/n_reads:8.1f} per read")
