#!/usr/bin/env python
"""Turn an ncu launch list into per-kernel / per-grid tables.
Input: the CSV of `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`
(one row per launch and metric).  With the DRAM metrics present the per-grid table also shows the bytes each launch
moved and the resulting GB/s (cold-cache, serialised launches: compare shares and traffic, not absolute times).
usage: python profiles/summarize.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import collections
import csv
import re
import sys


def to_us(v, unit):
    return {"ns": v / 1e3, "us": v, "usecond": v, "ms": v * 1e3, "msecond": v * 1e3, "s": v * 1e6}.get(unit, v / 1e3)


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    launches = collections.OrderedDict()          # ID -> dict
    for r in csv.DictReader(lines):
        L = launches.setdefault(r["ID"], {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", ""),
                                          "grid": r["Grid Size"], "block": r["Block Size"], "us": 0.0, "rd": None, "wr": None})
        v = float(r["Metric Value"].replace(",", ""))
        m = r["Metric Name"]
        if m.startswith("gpu__time_duration"):
            L["us"] = to_us(v, r["Metric Unit"])
        elif m.startswith("dram__bytes_read"):
            L["rd"] = to_bytes(v, r["Metric Unit"])
        elif m.startswith("dram__bytes_write"):
            L["wr"] = to_bytes(v, r["Metric Unit"])
    rows = list(launches.values())
    have_dram = any(r["rd"] is not None for r in rows)
    per_kernel = collections.defaultdict(lambda: [0, 0.0])
    per_grid = collections.defaultdict(lambda: [0, 0.0, 0.0])
    total = 0.0
    for r in rows:
        per_kernel[r["name"]][0] += 1
        per_kernel[r["name"]][1] += r["us"]
        g = per_grid[(r["name"], r["grid"], r["block"])]
        g[0] += 1
        g[1] += r["us"]
        g[2] += (r["rd"] or 0.0) + (r["wr"] or 0.0)
        total += r["us"]
    print(f"# launch list: {path}\n\n{len(rows)} kernel launches, {total:.1f} us summed (serialised, cold-cache: compare shares)\n")
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (n, t) in sorted(per_kernel.items(), key=lambda x: -x[1][1]):
        print(f"| {k} | {n} | {t:.1f} | {100 * t / total:.1f}% |")
    if have_dram:
        print("\n| kernel | grid | block | launches | avg us | total us | DRAM MB / launch | DRAM GB/s |\n|---|---|---|---:|---:|---:|---:|---:|")
    else:
        print("\n| kernel | grid | block | launches | avg us | total us |\n|---|---|---|---:|---:|---:|")
    for (k, g, b), (n, t, by) in sorted(per_grid.items(), key=lambda x: -x[1][1])[:32]:
        if have_dram:
            print(f"| {k} | {g} | {b} | {n} | {t / n:.2f} | {t:.1f} | {by / n / 1e6:.2f} | {by / t / 1e3 if t else 0:.0f} |")
        else:
            print(f"| {k} | {g} | {b} | {n} | {t / n:.2f} | {t:.1f} |")
    if have_dram:
        print("\nLargest launch of each kernel (the finest level it runs on):\n")
        print("| kernel | grid | us | DRAM read MB | DRAM written MB | DRAM GB/s |\n|---|---|---:|---:|---:|---:|")
        best = {}
        for r in rows:
            if r["name"] not in best or r["us"] > best[r["name"]]["us"]:
                best[r["name"]] = r
        for k, r in sorted(best.items(), key=lambda x: -x[1]["us"]):
            by = (r["rd"] or 0.0) + (r["wr"] or 0.0)
            print(f"| {k} | {r['grid']} | {r['us']:.2f} | {(r['rd'] or 0) / 1e6:.1f} | {(r['wr'] or 0) / 1e6:.1f} | {by / r['us'] / 1e3 if r['us'] else 0:.0f} |")


if __name__ == "__main__":
    main(sys.argv[1])
