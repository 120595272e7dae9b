#!/usr/bin/env python
"""Turn an ncu launch list (--metrics gpu__time_duration.sum --csv) into a per-kernel / per-grid table.
usage: python profiles/summarize.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    per_kernel = collections.defaultdict(lambda: [0, 0.0])
    per_grid = collections.defaultdict(lambda: [0, 0.0])
    total = 0.0
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        per_kernel[name][0] += 1; per_kernel[name][1] += us
        key = (name, r["Grid Size"], r["Block Size"])
        per_grid[key][0] += 1; per_grid[key][1] += us
        total += us
    print(f"# launch list: {path}\n\n{len(rows)} kernel launches, {total:.1f} us summed (serialised, cold-cache: compare shares)\n")
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (n, t) in sorted(per_kernel.items(), key=lambda x: -x[1][1]):
        print(f"| {k} | {n} | {t:.1f} | {100 * t / total:.1f}% |")
    print("\n| kernel | grid | block | launches | avg us | total us |\n|---|---|---|---:|---:|---:|")
    for (k, g, b), (n, t) in sorted(per_grid.items(), key=lambda x: -x[1][1])[:30]:
        print(f"| {k} | {g} | {b} | {n} | {t / n:.2f} | {t:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
