#!/usr/bin/env python
"""Per-source-line share of executed warp instructions and stall samples of an .ncu-rep captured with --import-source on
(the kernel must be compiled with -lineinfo).  usage: ncu_lines.py file.ncu-rep [top]"""
import collections
import csv
import subprocess
import sys


def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    inst, stall, text = collections.Counter(), collections.Counter(), {}
    cur, fname = None, ""
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if r[0].isdigit():
            cur = (fname, int(r[0]))
            text[cur] = ",".join(r[1:]).strip()[:110]
            continue
        if hdr and r[0] == "" and len(r) > 7 and r[2].startswith("0x"):
            try:
                inst[cur] += float(r[hdr.index("Instructions Executed")])
                stall[cur] += float(r[hdr.index("Warp Stall Sampling (All Samples)")])
            except ValueError:
                pass
    ti, ts = sum(inst.values()), sum(stall.values())
    print(f"{path}: {ti:.0f} warp instructions, {ts:.0f} stall samples")
    for key, n in inst.most_common(top):
        print(f"{100 * n / ti:5.1f}% inst {100 * stall[key] / max(ts, 1):5.1f}% stall  {key[0]}:{key[1]}  {text.get(key, '')}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
