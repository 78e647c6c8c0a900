"""Top stall lines of an `ncu --page source --csv` export (SASS view).  Usage: ncu_source_top.py file.csv [n]"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1]))]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
rows = rows[starts[which]:starts[which + 1]]
print(rows[0][1][:100])
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for idx, r in enumerate(rows[2:]):
    if len(r) < len(hdr):
        continue
    v = float(r[ci["# Samples"]] or 0)
    data.append((v, idx, r))
tot = sum(v for v, _, _ in data)
print("total samples %d, instructions %d" % (tot, len(data)))
agg = {s: sum(float(r[ci[s]] or 0) for _, _, r in data) for s in stalls}
print("stall totals:", ", ".join("%s=%.1f%%" % (k[6:], 100 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for v, idx, r in sorted(data, key=lambda x: -x[0])[:n]:
    top = sorted(((float(r[ci[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print("%7.0f %5.1f%%  #%-5d exec=%-9s %-70s %s" % (v, 100 * v / max(tot, 1), idx, r[ci["Instructions Executed"]], r[ci["Source"]].strip()[:70],
                                                " ".join("%s:%d" % (s, c) for c, s in top if c)))
