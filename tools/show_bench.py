"""Print the headline fields of bench.py JSON lines: python tools/show_bench.py file..."""
import json, sys
for f in sys.argv[1:]:
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        r = j.get("roofline", {})
        cb = j.get("cpu_baseline", {})
        print(f"{f}: {j['ms_per_step']:.4f} ms/step  {j['value']/1e6:.2f} M tok/s  roofline {r.get('achieved')} {r.get('unit')} frac {r.get('frac')}"
              f"  e2e {j['e2e']['ms_per_step']:.3f} ms ({j['e2e']['value']/1e6:.2f} M tok/s)  cpu {cb.get('value')} on {cb.get('cores')} cores"
              f"  exact={j.get('histogram_counts_exact')} clocks={j.get('clocks')}")
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-400:])
