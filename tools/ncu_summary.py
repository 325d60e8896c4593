#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total / mean time and
share of the captured GPU time.  Usage: python tools/ncu_summary.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import csv
import io
import re
import sys
from collections import OrderedDict


def main(path):
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = OrderedDict()
    total = 0.0
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit", "ns") in ("us", "usecond"):
            ns *= 1e3
        elif r.get("Metric Unit", "ns") in ("ms", "msecond"):
            ns *= 1e6
        name = re.sub(r"\(.*$", "", r["Kernel Name"])
        a = agg.setdefault(name, dict(n=0, ns=0.0, grid=r["Grid Size"], block=r["Block Size"]))
        a["n"] += 1
        a["ns"] += ns
        total += ns
    print(f"source: `{path}` — {sum(a['n'] for a in agg.values())} launches, {total / 1e6:.2f} ms of GPU time captured "
          "(ncu serialises launches and runs them cold-cache: compare SHARES with bench.py's `kernels`, not absolute times)\n")
    print("| kernel | launches | total ms | mean µs | share | grid (first) | block |")
    print("|---|---:|---:|---:|---:|---|---|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        print(f"| `{name}` | {a['n']} | {a['ns'] / 1e6:.3f} | {a['ns'] / a['n'] / 1e3:.1f} | {100 * a['ns'] / total:.1f} % | {a['grid']} | {a['block']} |")


if __name__ == "__main__":
    main(sys.argv[1])
