"""dev helper: one-screen summary of a bench.py JSON line (and of the reference arm's)"""
import json, sys
j = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])  # (an NCCL banner may precede the line)
if j.get("impl") == "reference":
    for c in j["configs"]:
        print(c["name"], round(c.get("value", 0)), c.get("unavailable", ""), c.get("cpu_baseline", {}).get("cores"))
    sys.exit(0)
print("headline value %.1fM e2e %.1fM e2e_lists %.1fM e2e_ascii %.1fM launches %s" % (j['value'] / 1e6, j['e2e']['value'] / 1e6, j['e2e_lists']['value'] / 1e6, j['e2e_ascii']['value'] / 1e6, j['gpu_launches']))
print("clocks", j['clocks'])
for c in j['configs']:
    if 'unavailable' in c:
        print(c)
        continue
    print(c['name'], "value %.1fM  e2e %.1fM  lists %.1fM  ascii %.1fM" % (c['value'] / 1e6, c['e2e']['value'] / 1e6, c['e2e_lists']['value'] / 1e6, c['e2e_ascii']['value'] / 1e6),
          "kernel_ms", {k: round(v, 2) for k, v in c['kernel_ms'].items()}, "cpu", c['cpu_baseline'] and (round(c['cpu_baseline']['value']), c['cpu_baseline']['matches_gpu_output']),
          "colors/read %.1f" % (c['results_total_colors'] / c['reads_per_gpu']))
    r = c['roofline']
    print("    roofline", r['bound'], "frac %.3f" % r['frac'], r['kernel'], "| hbm models: indep %.0f B/read frac %.2f ; seed %.0f B/read frac %.3f seeds/read %.1f" % (
        r['hbm']['algorithmic_model']['bytes_per_read'], r['hbm']['algorithmic_model']['frac'], r['hbm']['seed_extend_model']['bytes_per_read'],
        r['hbm']['seed_extend_model']['frac'], r['hbm']['seed_extend_model']['seeds_per_read']))
print("e2e_tool", j.get('e2e_tool'))
