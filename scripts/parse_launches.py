"""Summarises an ncu launch list (csv[.gz] of --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]) per kernel."""
import collections
import csv
import gzip
import io
import sys


def load(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    launch = collections.OrderedDict()
    for x in csv.DictReader(io.StringIO("".join(lines))):
        d = launch.setdefault(int(x["ID"]), {"name": x["Kernel Name"], "grid": x["Grid Size"]})
        d[x["Metric Name"]] = float(x["Metric Value"].replace(",", ""))
    return launch


def short(name):
    n = name.replace("b2::", "").split("(")[0]
    return n.replace("void ", "")


if __name__ == "__main__":
    launch = load(sys.argv[1])
    per = collections.OrderedDict()
    for k, v in launch.items():
        p = per.setdefault(short(v["name"]), [0, 0.0, 0.0, 0.0])
        p[0] += 1
        p[1] += v.get("gpu__time_duration.sum", 0.0) / 1e6
        p[2] += v.get("dram__bytes_read.sum", 0.0) / 1e9
        p[3] += v.get("dram__bytes_write.sum", 0.0) / 1e9
    tot = sum(p[1] for p in per.values())
    print("| kernel | launches | total ms | share | DRAM read GB | DRAM write GB |\n|---|---|---|---|---|---|")
    for n, p in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print(f"| {n} | {p[0]} | {p[1]:.2f} | {p[1] / tot:.3f} | {p[2]:.2f} | {p[3]:.2f} |")
    print(f"\ntotal {tot:.1f} ms, DRAM {sum(p[2] + p[3] for p in per.values()):.1f} GB")
    if len(sys.argv) > 2:
        for k, v in launch.items():
            print(k, short(v["name"]), v["grid"], round(v.get("gpu__time_duration.sum", 0) / 1e6, 3), "ms", round((v.get("dram__bytes_read.sum", 0) + v.get("dram__bytes_write.sum", 0)) / 1e9, 2), "GB")
