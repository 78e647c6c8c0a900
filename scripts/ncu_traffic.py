"""profiles/ncu_traffic.json from an `ncu --page raw --csv` export: per kernel (bench.py's names) the DRAM
bytes (read + write) of its LARGEST launch -- the `roofline.traffic` value bench.py reports."""
import csv
import json
import sys

MAP = [("pad3d_kernel", "pad3d_kernel"), ("spline_up_z_kernel", "spline_up_z_kernel"),
       ("spline_up_strided_kernel<double, double", "spline_up_x_kernel"), ("spline_up_strided_kernel<double, float", "spline_up_y_kernel"),
       ("log_pass_strided_kernel", None), ("log_pass_z_kernel", "log_pass_z_kernel"), ("log_pass_yz_kernel", "log_pass_yz_kernel"), ("gradient_masked_kernel", "gradient_masked_kernel"), ("gradient_kernel", "gradient_kernel"),
       ("detect_peaks_kernel", "detect_peaks_kernel"), ("orient_kernel", "orient_kernel"), ("describe_kernel", "describe_kernel"),
       ("match_u8_kernel<0", "match_u8_pairs_kernel"), ("match_u8_kernel<2", "match_u8_topk_kernel")]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def mb(r, key):
    v = float(r[col[key]].replace(",", ""))
    u = units[col[key]].lower()
    return v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)


out = {}
for r in data:
    name = r[col["Kernel Name"]]
    for pat, short in MAP:
        if pat in name:
            if short is None:      # log_pass_strided: MODE 0 = x pass, MODE 1 = y pass (last template argument)
                short = "log_pass_y_kernel" if ", 1>" in name.split("(")[0] else "log_pass_x_kernel"
            b = mb(r, "dram__bytes_read.sum") + mb(r, "dram__bytes_write.sum")
            out[short] = max(out.get(short, 0), int(b))
            break
json.dump(out, open(sys.argv[2], "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1, sort_keys=True))
