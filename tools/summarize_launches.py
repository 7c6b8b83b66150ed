#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launch count, total time, share.

    python tools/summarize_launches.py gpurun_out/<dir>/launches.csv > profiles/<name>.md
"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hdr]
    ki, vi, gi, bi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        a = agg.setdefault(name, dict(n=0, ns=0.0, grid=r[gi], block=r[bi]))
        a["n"] += 1
        a["ns"] += float(r[vi].replace(",", ""))
    tot = sum(a["ns"] for a in agg.values())
    print(f"source: {path}  ({sum(a['n'] for a in agg.values())} launches, {tot / 1e6:.3f} ms of kernel time; per-launch "
          "times are cold-cache and serialised under ncu: read the SHARE, not the absolute)\n")
    print("| kernel | launches | grid | block | total ms | avg us | share |")
    print("|---|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        print(f"| `{k}` | {a['n']} | {a['grid']} | {a['block']} | {a['ns'] / 1e6:.3f} | {a['ns'] / a['n'] / 1e3:.1f} | {a['ns'] / tot:.3f} |")


if __name__ == "__main__":
    main(sys.argv[1])
