"""dev: condense an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into one step
(the launches between the last two fps_kernel launches): per-kernel totals, shares and the ordered list of launches >= 50 us.
usage: python tests/tools/launch_summary.py launches.csv [out_summary.txt]"""
import csv
import re
import sys


def short(n):
    n = re.sub(r'^void ', '', n).replace('vgtkb::', '')
    m = re.match(r'([\w:<>, ]+?)\(', n)
    return (m.group(1) if m else n)[:70]


rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
fps = [i for i, r in enumerate(rows) if 'fps_kernel' in r[4]]
assert len(fps) >= 2, "need two fps_kernel launches to delimit a step"
step = rows[fps[-2]:fps[-1]]
tot = {}
for r in step:
    d = tot.setdefault(short(r[4]), [0.0, 0])
    d[0] += float(r[-1]) / 1e3
    d[1] += 1
total = sum(v[0] for v in tot.values())
out = [f"one step of bench.py (launches between two fps_kernel launches): {len(step)} launches, {total / 1e3:.2f} ms summed "
       "(cold-cache, serialised under ncu: compare shares)"]
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    out.append(f"{k:72s} {v[0]:8.1f} us x{v[1]:3d} {100 * v[0] / total:5.1f}%")
out.append("")
out.append("launches >= 50 us in order:")
for r in step:
    t = float(r[-1]) / 1e3
    if t >= 50:
        out.append(f"  {short(r[4]):60s} grid={r[8]:16s} {t:8.1f} us")
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
