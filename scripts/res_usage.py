"""profiles/r02_res_usage.txt: registers per thread, stack, static shared memory and local memory of every kernel of
libmad_b200.so (cuobjdump -res-usage).  Runs without a GPU.  usage: python scripts/res_usage.py [lib] > profiles/..."""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "mad_b200/libmad_b200.so"
txt = subprocess.run(["cuobjdump", "-res-usage", lib], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows, name = [], None
for ln in txt.splitlines():
    m = re.match(r"\s*Function (\S+):", ln)
    if m:
        name = m.group(1)
        continue
    m = re.match(r"\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", ln)
    if m and name:
        rows.append((name,) + tuple(int(x) for x in m.groups()))
        name = None
dem = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), stdout=subprocess.PIPE, text=True).stdout.splitlines()
print("# cuobjdump -res-usage of %s (sm_100a): registers per thread, stack bytes, static shared bytes, local bytes" % lib)
for r, d in zip(rows, dem):
    d = re.sub(r"\(anonymous namespace\)::", "", d).split("(")[0][:90]
    print("%-92s REG %3d  STACK %5d  SHARED %6d  LOCAL %4d" % (d, r[1], r[2], r[3], r[4]))
