"""dev: compact view of a bench.py JSON line (value, e2e, per-entry-point ms and roofline fractions)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "clocks", d.get("clocks"))
r = d["roofline"]
print("roofline", r["kernel"], r["bound"], round(r["achieved"], 1), r["unit"], "frac", round(r["frac"], 3))
oth = d.get("rooflines_other", {})
for k, v in d["kernel_table"].items():
    o = oth.get(k)
    print(f"  {k:34s} {v['ms_per_step']:7.3f} ms  x{v['calls_per_step']:.0f}" + (f"   {o['bound']} {o['achieved']:.0f} {o['unit']} frac {o['frac']:.3f}" if o else ""))
