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
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in csv.DictReader(rows):
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0, 0.0])
        if r["Metric Name"] == "gpu__time_duration.sum":
            a[0] += 1
            a[1] += float(r["Metric Value"].replace(",", "")) / 1e3
        elif r["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            a[2] += float(r["Metric Value"].replace(",", "")) * scale.get(r["Metric Unit"], 1.0)
    total = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {total / 1e3:.2f} ms device time (ncu: serialised, cold cache)")
    print(f"{'share':>7} {'total_us':>11} {'launches':>8} {'avg_us':>9} {'dram_MB/launch':>14}  kernel")
    for k, (n, us, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{us / total:7.3%} {us:11.1f} {n:8d} {us / n:9.2f} {by / n / 1e6:14.2f}  {k}")
    return agg


if __name__ == "__main__":
    agg = main(sys.argv[1])
    if len(sys.argv) > 2:  # second argument: write profiles/traffic.json (dram bytes per launch of the library's kernels)
        import json
        entry = {"gemm_bf16_tma": "fmc_gemm_bf16", "spatial_attn": "fmc_spatial_attn_bf16", "groupnorm": "fmc_groupnorm_bf16",
                 "layernorm": "fmc_layernorm_bf16", "temporal_qkv_attn": "fmc_temporal_qkv_attn_bf16",
                 "conv3x3": "fmc_conv3x3_bf16"}
        out = {}
        for k, (n, us, by) in agg.items():
            for frag, name in entry.items():
                if frag in k and by > 0:
                    o = out.setdefault(name, {"launches": 0, "bytes": 0.0})
                    o["launches"] += n
                    o["bytes"] += by
        res = {name: round(o["bytes"] / o["launches"]) for name, o in out.items()}
        res["_source"] = (f"{sys.argv[1]}: mean dram__bytes_read.sum + dram__bytes_write.sum per launch over every launch of the "
                          "kernel in two bench steps, ncu --cache-control all (cold L2 before every kernel: an upper bound "
                          "of the traffic inside the running step)")
        json.dump(res, open(sys.argv[2], "w"), indent=1)
