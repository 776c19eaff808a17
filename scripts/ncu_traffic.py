"""profiles/ncu_traffic.json from raw-page CSV exports of `ncu --set full` captures (bench.py reads it for
roofline.traffic):   python scripts/ncu_traffic.py gpurun_out/prof_A_raw.csv [gpurun_out/prof_B_raw.csv ...]"""
import csv
import json
import os
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {"note": "dram__bytes_read.sum / dram__bytes_write.sum per launch from ncu --set full captures of "
               "scripts/ncu_ops.py (2^28 fp32 elements; GEMM 4096^3 and 8192^3)", "kernels": {}}
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    h, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(h)}
    for r in rows[2:]:
        name = r[col["Kernel Name"]].replace("void ", "").split("(")[0].replace("tc::", "")
        grid = r[col["launch__grid_size"]]
        if name in out["kernels"] and "gemm" not in name:
            continue   # first launch of each kernel (the bench-size one)
        key = name if "gemm" not in name else f"{name} grid={grid}"
        rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * UNIT[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * UNIT[units[col["dram__bytes_write.sum"]]]
        out["kernels"][key] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "source": os.path.basename(path)}
json.dump(out, open("profiles/ncu_traffic.json", "w"), indent=1)
print(len(out["kernels"]), "kernels")
