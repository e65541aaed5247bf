"""Per-kernel summary of an `ncu --page raw --csv` export: launches, mean duration, DRAM bytes per launch, achieved DRAM GB/s against the
measured HBM copy peak (MEASURED_PEAKS.json), L2 hit rate, registers.  python scripts/parse_ncu_raw.py file.csv[.gz]"""
import collections
import csv
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def f(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def main(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as fh:
        rows = list(csv.reader(fh))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6529.4
    scale_t = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[units[ix["gpu__time_duration.sum"]]]
    scale_b = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    agg = collections.OrderedDict()
    for r in data:
        name = r[ix["Kernel Name"]].split("(")[0].replace("b2::", "").replace("void ", "")
        t = f(r[ix["gpu__time_duration.sum"]]) * scale_t
        rd = f(r[ix["dram__bytes_read.sum"]]) * scale_b[units[ix["dram__bytes_read.sum"]]]
        wr = f(r[ix["dram__bytes_write.sum"]]) * scale_b[units[ix["dram__bytes_write.sum"]]]
        a = agg.setdefault(name, {"n": 0, "t": 0.0, "b": 0.0, "best": 0.0, "l2": [], "regs": r[ix["launch__registers_per_thread"]], "warps": []})
        a["n"] += 1
        a["t"] += t
        a["b"] += rd + wr
        a["best"] = max(a["best"], (rd + wr) / t / 1e9 if t > 0 else 0.0)
        a["l2"].append(f(r[ix["lts__t_sector_hit_rate.pct"]]))
        a["warps"].append(f(r[ix["sm__warps_active.avg.pct_of_peak_sustained_active"]]))
    print(f"| kernel | launches | mean µs | DRAM MB per launch | mean GB/s | best GB/s | best / measured HBM peak ({peak:.0f} GB/s) | L2 hit % | regs | warps active % |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for n, a in agg.items():
        print(f"| {n} | {a['n']} | {a['t'] / a['n'] * 1e6:.1f} | {a['b'] / a['n'] / 1e6:.1f} | {a['b'] / a['t'] / 1e9:.0f} | {a['best']:.0f} | {a['best'] / peak:.2f} | "
              f"{sum(a['l2']) / len(a['l2']):.0f} | {a['regs']} | {sum(a['warps']) / len(a['warps']):.0f} |")


if __name__ == "__main__":
    main(sys.argv[1])
