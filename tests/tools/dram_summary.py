"""dev: condense an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` capture of
tests/tools/one_step.py into the per-kernel totals of ONE step (the launches between the last two fps_kernel launches):
launches, summed time, DRAM bytes read / written, achieved DRAM rate.
usage: python tests/tools/dram_summary.py capture.csv [out.txt]"""
import csv
import re
import sys


def short(n):
    n = re.sub(r'^void ', '', n).replace('vgtkb::', '')
    m = re.match(r'([\w:<>, ]+?)\(', n)
    return (m.group(1) if m else n)[:64]


rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
launch = {}
order = []
for r in rows:
    i = int(r[0])
    if i not in launch:
        launch[i] = {"name": r[4]}
        order.append(i)
    v = float(r[-1].replace(',', ''))
    unit = r[-2]
    if unit in ("Kbyte", "Mbyte", "Gbyte"):
        v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    if unit in ("us", "usecond"):
        v *= 1e3
    if unit in ("ms", "msecond"):
        v *= 1e6
    launch[i][r[-3]] = v
fps = [k for k, i in enumerate(order) if 'fps_kernel' in launch[i]["name"]]
assert len(fps) >= 2, "need two fps_kernel launches to delimit a step"
step = [launch[i] for i in order[fps[-2]:fps[-1]]]
tot = {}
for L in step:
    d = tot.setdefault(short(L["name"]), [0, 0.0, 0.0, 0.0])
    d[0] += 1
    d[1] += L.get("gpu__time_duration.sum", 0.0) / 1e3
    d[2] += L.get("dram__bytes_read.sum", 0.0)
    d[3] += L.get("dram__bytes_write.sum", 0.0)
T = sum(v[1] for v in tot.values())
R = sum(v[2] for v in tot.values())
W = sum(v[3] for v in tot.values())
out = [f"one step ({len(step)} launches): {T / 1e3:.2f} ms summed under ncu (serialised, cold caches), DRAM read {R / 1e9:.2f} GB, "
       f"written {W / 1e9:.2f} GB -> {(R + W) / 1e9:.2f} GB per step",
       f"{'kernel':66s} {'n':>3s} {'us':>9s} {'read MB':>9s} {'write MB':>9s} {'GB/s':>7s}"]
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    rate = (v[2] + v[3]) / (v[1] * 1e-6) / 1e9 if v[1] > 0 else 0.0
    out.append(f"{k:66s} {v[0]:3d} {v[1]:9.1f} {v[2] / 1e6:9.1f} {v[3] / 1e6:9.1f} {rate:7.0f}")
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
