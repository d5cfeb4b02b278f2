"""dev helper: warp instructions per source REGION (line ranges given as name:file:lo-hi ...) from an ncu source-page CSV.
usage: ncu_regions.py src.csv n_reads name:file:lo-hi ..."""
import csv, sys, collections
rows = csv.reader(open(sys.argv[1])); n_reads = float(sys.argv[2])
regions = []
for a in sys.argv[3:]:
    name, f, rng = a.split(':'); lo, hi = rng.split('-'); regions.append((name, f, int(lo), int(hi)))
agg = collections.Counter(); samp = collections.Counter(); cur = None; hdr = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; ii = r.index('Instructions Executed'); isamp = r.index('# Samples'); continue
    if hdr is None or len(r) < 8 or r[0] == '' or r[2] != '-': continue
    try: ln = int(r[0]); inst = int(r[ii]); sm = int(r[isamp])
    except ValueError: continue
    key = 'other:' + cur
    for name, f, lo, hi in regions:
        if f == cur and lo <= ln <= hi: key = name; break
    agg[key] += inst; samp[key] += sm
tot = sum(agg.values()); ts = sum(samp.values())
for k, v in agg.most_common(): print(f"{k:40s} {v/n_reads:8.1f} inst/read {100*v/tot:5.1f}%   samples {100*samp[k]/ts:5.1f}%")
print("total", tot / n_reads)
