#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total device time, share.
usage: python profiles/summarise_launches.py gpurun_out/launches.csv > profiles/rNN_launches.txt"""
import csv
import re
import sys


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    return name[:90]


def main(path):
    rows = [l for l in open(path) if l.startswith('"')]
    agg = {}
    for r in csv.DictReader(rows):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"]) / 1e3
    total = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {total / 1e3:.2f} ms device time (ncu: serialised, cold cache)")
    print(f"{'share':>7} {'total_us':>11} {'launches':>8} {'avg_us':>9}  kernel")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{us / total:7.3%} {us:11.1f} {n:8d} {us / n:9.2f}  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
