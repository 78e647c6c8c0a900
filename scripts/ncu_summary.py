"""Summarise an `ncu --page raw --csv` export: one line per kernel launch with the counters the
roofline discussion needs.  Usage: python scripts/ncu_summary.py raw.csv > profiles/xxx.txt"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
units = rows[1]
data = rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = [
    ("gpu__time_duration.sum", "dur_us", 1e-3),
    ("dram__bytes_read.sum", "rd_MB", None),
    ("dram__bytes_write.sum", "wr_MB", None),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%", 1),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%", 1),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%", 1),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%", 1),
]


def to_mb(v, u):
    v = float(v.replace(",", ""))
    u = u.lower()
    return v * {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1e-6)


print("%-46s %9s %9s %9s %6s %6s %6s %6s %7s %6s %6s %5s %6s  grid" % (
    ("kernel",) + tuple(w[1] for w in want)))
agg = {}
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    if len(name) > 46:
        name = name[:46]
    vals = []
    for key, label, scale in want:
        if key not in col or r[col[key]] == "":
            vals.append(float("nan"))
            continue
        raw = r[col[key]]
        if scale is None:
            vals.append(to_mb(raw, units[col[key]]))
        else:
            v = float(raw.replace(",", ""))
            if label == "dur_us":
                u = units[col[key]].lower()
                v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(u, 1e-3)
            vals.append(v)
    grid = r[col["Grid Size"]] if "Grid Size" in col else ""
    print("%-46s %9.1f %9.1f %9.1f %6.1f %6.1f %6.1f %6.1f %7.1f %6.1f %6.1f %5.0f %6.1f  %s" % ((name,) + tuple(vals) + (grid,)))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += vals[0]
print()
tot = sum(v[1] for v in agg.values())
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-46s launches %3d  total %9.1f us  share %5.1f%%" % (name, n, t, 100 * t / tot))
