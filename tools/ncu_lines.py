"""dev helper: per-source-line instruction and stall-sample totals from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None; hdr = None
agg = collections.OrderedDict()
kernel_seen = 0
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name':
        kernel_seen += 1; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None: continue
    if r[0] != '' and r[2] == '-':  # source line row
        try:
            key = (cur_file, int(r[0]), r[1].strip()[:110])
            inst = int(r[hdr.index('Instructions Executed')]); samp = int(r[hdr.index('# Samples')])
        except ValueError:
            continue
        a = agg.setdefault(key, [0, 0]); a[0] += inst; a[1] += samp
tot_i = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
print("total inst", tot_i, "samples", tot_s)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*v[0]/tot_i:5.1f}% inst {100*v[1]/max(1,tot_s):5.1f}% samp  {k[0]}:{k[1]}  {k[2]}")
