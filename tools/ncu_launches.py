#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share.

    python tools/ncu_launches.py gpurun_out/launches.csv
"""
import collections
import csv
import sys


def main(path):
    lines = open(path, errors="replace").read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(lines[start:]))
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = r["Kernel Name"].split("(")[0].replace("void ", "")[:70]
        v = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit") == "us":
            v *= 1e3
        tot[k] += v
        cnt[k] += 1
    s = sum(tot.values())
    print(f"{len(rows)} launches, {s / 1e3:.1f} us of device time (cold-cache, serialised under ncu: compare shares)")
    for k, v in tot.most_common(12):
        print(f"{k:72s} n={cnt[k]:4d} total {v / 1e3:10.1f} us {100 * v / s:5.1f}%  avg {v / cnt[k] / 1e3:8.1f} us")


if __name__ == "__main__":
    main(sys.argv[1])
