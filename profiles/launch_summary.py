#!/usr/bin/env python
"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X) per kernel.
Usage: launch_summary.py launches.csv [steps]   (steps: number of MD steps the captured window covers)"""
import collections
import csv
import sys


def main(path, steps=None):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    if not rows:
        print("no kernels in", path)
        return
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0, []])
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        a = agg[r[ki].split("(")[0][:70]]
        a[0] += 1
        a[1] += v
        a[2].append(v)
    total = sum(t for _, t, _ in agg.values())
    print(f"{len(rows) - 1} launches, {total:.1f} us of kernel time" + (f", {total / steps:.1f} us per step" if steps else ""))
    print("(mean includes the launches queued behind a rebuild decision, which return at once: read the median for a kernel's duration)")
    for k, (c, t, vs) in sorted(agg.items(), key=lambda x: -x[1][1]):
        per = f" {t / steps:8.1f} us/step" if steps else ""
        med = sorted(vs)[len(vs) // 2]
        print(f"{t:10.1f} us {100 * t / total:5.1f}% {c:5d} x mean {t / c:8.1f} median {med:8.1f} us{per}  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
