"""dev helper: (re)writes one entry of profiles/kernels.json from the raw-metrics page of an `ncu --set full` capture.
usage: ncu_kernels_json.py KERNEL@INDEX READS_IN_LAUNCH RAW.csv [SOURCE-NOTE]
bench.py reads profiles/kernels.json for roofline.frac (instruction issue) and roofline.traffic (DRAM bytes)."""
import csv, json, os, sys

key, reads, raw = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[-1]
col = {h: i for i, h in enumerate(hdr)}


def f(name, default=None):
    i = col.get(name)
    if i is None or vals[i] == "":
        return default
    v = float(vals[i].replace(",", ""))
    u = units[i]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1e-3, "ms": 1.0, "s": 1e3, "ns": 1e-6}.get(u, 1.0)


stalls = {}
for h, i in col.items():
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        n = h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]
        try:
            v = float(vals[i])
        except ValueError:
            continue
        if v >= 0.8 and n not in ("selected", "not_selected"):
            stalls[n] = round(v, 2)
entry = {
    "reads_in_launch": reads,
    "ms": round(f("gpu__time_duration.sum"), 4),
    "warp_instructions_per_read": round(f("smsp__inst_executed.sum") / reads, 1),
    "dram_bytes_per_read": round((f("dram__bytes_read.sum") + f("dram__bytes_write.sum")) / reads, 1),
    "l2_hit_pct": round(f("lts__t_sector_hit_rate.pct"), 1),
    "alu_pipe_pct_of_peak": round(f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), 1),
    "issue_slots_pct_of_peak": round(f("smsp__issue_active.avg.pct_of_peak_sustained_active"), 1),
    "warps_active_pct_of_peak": round(f("sm__warps_active.avg.pct_of_peak_sustained_active"), 1),
    "registers_per_thread": int(f("launch__registers_per_thread")),
    "dram_throughput_pct": round(f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0.0), 1),
    "l2_throughput_pct": round(f("lts__throughput.avg.pct_of_peak_sustained_elapsed", 0.0), 1),
    "top_stalls_per_issue": dict(sorted(stalls.items())),
    "source": (sys.argv[4] if len(sys.argv) > 4 else "profiles/" + os.path.basename(raw)) + " (ncu --set full --clock-control none, bench.py --kernel-only)",
}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "kernels.json")
d = json.load(open(path))
d[key] = entry
json.dump(d, open(path, "w"), indent=1)
print(key, json.dumps(entry))
