"""Summarise an .ncu-rep (read on the CPU box) into a small text table for profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_summary.txt
"""
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "dur_us"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def main():
    rep = sys.argv[1]
    if rep.endswith(".csv"):   # a raw page already exported on the GPU box (ncu -i X.ncu-rep --page raw --csv)
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(h)}
    print(f"# {rep}: one row per captured launch (ncu --set full --clock-control none; cold-cache, serialised)")
    print("# " + " | ".join(["kernel"] + [f"{short}[{units[col[m]]}]" for m, short in METRICS if m in col]))
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        name = name.replace("void ", "").split("(")[0][:58]
        vals = []
        for m, _ in METRICS:
            if m in col:
                v = r[col[m]]
                try:
                    v = f"{float(v.replace(',', '')):.4g}"
                except ValueError:
                    pass
                vals.append(v)
        print(f"{name:58s} | " + " | ".join(vals))


if __name__ == "__main__":
    main()
