"""dev: instruction counts of the shipped library (cuobjdump -sass): the Blackwell-native claim in numbers.
usage: python tests/tools/sass_counts.py [libvgtkb200.so] > profiles/r2_sass_counts.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "equi_articulated_pose_b200/libvgtkb200.so"
ops = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "HMMA", "REDUX", "ELECT", "R2UR", "SYNCS",
       "UTMACMDFLUSH", "UTMAPF"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
per, total, cur, k = collections.OrderedDict(), collections.Counter(), None, 0
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = re.sub(r"\(.*", "", names[k])
        k += 1
        per[cur] = collections.Counter()
        continue
    m = re.search(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        op = m.group(1)
        for o in ops:
            if op == o or op.startswith(o + "."):
                per[cur][o] += 1
                total[o] += 1
print(f"SASS of {lib} (cuobjdump -sass, sm_100a), instruction counts.")
print("UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = cp.async.bulk.tensor load / store, UBLKCP = cp.async.bulk,")
print("HMMA = warp-level mma.sync (grouping kernels), REDUX = redux.sync (FPS), ELECT = elect.sync, R2UR = vector -> uniform register move.")
print()
print("whole library: " + ", ".join(f"{o} {total[o]}" for o in ops))
print()
print("per kernel (only kernels with tensor-core / TMA / redux instructions):")
for name, c in per.items():
    if any(c[o] for o in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "REDUX")):
        print(f"  {name[:82]:82s} " + " ".join(f"{o}={c[o]}" for o in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "R2UR", "ELECT", "REDUX") if c[o]))
