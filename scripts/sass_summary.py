"""profiles/r02_sass_summary.txt: per kernel of libmad_b200.so the SASS opcode counts that prove the execution model
(cuobjdump -sass): tcgen05 MMAs (UTCIMMA / UTCHMMA), TMEM loads (LDTM), TMA (UTMALDG), cp.async (LDGSTS), FP64 (DFMA / DADD /
DMUL), conversions (F2F / I2F), shared-memory atomics (ATOMS), votes / match (VOTE / MATCH), local-memory spills (LDL / STL)."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "mad_b200/libmad_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
WANT = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG", "UTCBAR", "LDGSTS", "DFMA", "DADD", "DMUL", "F2F", "I2F", "FFMA", "ATOMS", "ATOMG", "RED",
        "VOTE", "MATCH", "REDUX", "SHFL", "LDL", "STL", "BAR", "SYNCS", "MEMBAR", "NANOSLEEP"]
cur, counts, total = None, collections.OrderedDict(), {}
for ln in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        total[cur] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for w in WANT:
            if op.startswith(w):
                counts[cur][w] += 1
                break
demangle = subprocess.run(["c++filt"], input="\n".join(counts), stdout=subprocess.PIPE, text=True).stdout.splitlines()
print("# SASS opcode counts per kernel of %s (cuobjdump -sass, sm_100a).  total = all instructions." % lib)
for name, pretty in zip(counts, demangle):
    pretty = re.sub(r"\(anonymous namespace\)::", "", pretty)
    pretty = pretty.split("(")[0][:90]
    c = counts[name]
    if total[name] < 20:
        continue
    print("%-92s total %6d  %s" % (pretty, total[name], "  ".join("%s %d" % (k, c[k]) for k in WANT if c[k])))
