#!/usr/bin/env python
"""Writes profiles/kernel_ncu.json (what bench.py's roofline block quotes) and the per-kernel summaries
profiles/r02_<name>_ncu.txt from the `ncu --set full` reports of profiles/ncu_capture.sh.
usage: python profiles/make_kernel_ncu.py gpurun_out"""
import csv
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CAPTURES = {  # report name -> (kernel name in bench.py, atoms the capture ran with)
    "lj_force": ("ljForceTiledKernel", 1000000),
    "build": ("verletBuildTiledKernel", 1000000),
    "adress_force": ("adressForceTiledKernel", 8000000),
    "molecule_force": ("moleculeForceTiledKernel", 16384000),
    "integrate": ("integratePreKernel", 1000000),
}


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, row = rows[0], rows[1], rows[2]
    return {h: (row[i], units[i]) for i, h in enumerate(hdr)}


def num(m, key):
    v, u = m[key]
    v = float(v.replace(",", ""))
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "%": 0.01}.get(u, 1.0)
    return v * scale


def main(src):
    table = {}
    for name, (kernel, atoms) in CAPTURES.items():
        path = os.path.join(src, f"r02_{name}.ncu-rep")
        if not os.path.exists(path):
            continue
        m = raw(path)
        entry = {
            "kernel": m["Kernel Name"][0][:120],
            "source": f"profiles/r02_{name}_ncu.txt (ncu --set full --clock-control none, first launch of the bench's timed region)",
            "duration_us_under_ncu": num(m, "gpu__time_duration.sum") * 1e6,
            "dram_bytes_read": num(m, "dram__bytes_read.sum"),
            "dram_bytes_write": num(m, "dram__bytes_write.sum"),
            "fp64_pipe_frac": num(m, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "issue_active_frac": num(m, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "dram_throughput_frac": num(m, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "l1tex_throughput_frac": num(m, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
            "warp_instructions": num(m, "smsp__inst_executed.sum"),
            "registers_per_thread": num(m, "launch__registers_per_thread"),
        }
        entry["dram_bytes_per_launch"] = entry["dram_bytes_read"] + entry["dram_bytes_write"]
        table[f"{kernel}@{atoms}"] = entry
        summary = subprocess.run([sys.executable, os.path.join(HERE, "ncu_summary.py"), path], capture_output=True, text=True).stdout
        open(os.path.join(HERE, f"r02_{name}_ncu.txt"), "w").write(summary)
        print(name, json.dumps({k: v for k, v in entry.items() if k not in ("kernel", "source")}))
    json.dump(table, open(os.path.join(HERE, "kernel_ncu.json"), "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out")
