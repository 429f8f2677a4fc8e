#!/usr/bin/env python
"""DRAM traffic per launch of the kernels in an `ncu --set full` capture -> profiles/r02_ncu_traffic.json, the file
bench.py's `roofline.traffic` reads (no hard-coded constant).

    python tools/ncu_traffic.py gpurun_out/<capture>.ncu-rep [more.ncu-rep ...] [--units Kernel=N ...] [--min-ms T]

--units Decompress=2097152 records how many units (points, additions) ONE captured launch processed, so that bench.py can
scale the measured bytes to the size of its own launches (bytes per unit x its units per launch).

For every captured launch: dram__bytes_read.sum + dram__bytes_write.sum, its duration, the SM-busy / pipe metrics that
the DESIGN's layout decision cites (LSU utilisation, long-scoreboard stalls), keyed by the functor name
(k_each<cpg::Decompress, ...> -> "Decompress").  Several launches of one kernel are averaged.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
WANT = {
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "gpu__time_duration.sum": "duration_ns",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active": "fmaheavy_pipe_pct",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed": "fmaheavy_cycles_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard_per_issue",
    "launch__registers_per_thread": "registers_per_thread",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
}
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e3, "us": 1e3, "msecond": 1e6, "ms": 1e6, "nsecond": 1.0, "ns": 1.0, "second": 1e9, "s": 1e9}


def short_name(name):
    m = re.search(r"k_each<(?:cpg::|\(anonymous namespace\)::|<unnamed>::)?(\w+)", name)
    if m:
        return m.group(1)
    return re.sub(r"\(.*", "", name).split("::")[-1]


def read(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name-base", "demangled"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        ent = {"kernel": short_name(r[col["Kernel Name"]])}
        for metric, key in WANT.items():
            if metric in col and r[col[metric]] not in ("", "n/a"):
                v = float(r[col[metric]].replace(",", ""))
                ent[key] = v * UNIT.get(units[col[metric]], 1.0)
        out.append(ent)
    return out


def main():
    args, units, min_ns = [], {}, 0.0
    it = iter(sys.argv[1:])
    for a in it:
        if a == "--min-ms":                  # leave out launches shorter than this (the 1-point self-test of cpg_init)
            min_ns = float(next(it)) * 1e6
        elif a == "--units":
            k, v = next(it).split("=")
            units[k] = float(v)
        else:
            args.append(a)
    launches = []
    for rep in args:
        for ent in read(rep):
            if ent.get("duration_ns", 0.0) < min_ns:
                continue
            ent["capture"] = os.path.basename(rep)
            launches.append(ent)
    agg = {}
    for ent in launches:
        k = ent["kernel"]
        a = agg.setdefault(k, {"launches": 0, "capture": ent["capture"]})
        a["launches"] += 1
        for key, v in ent.items():
            if isinstance(v, float):
                a[key] = a.get(key, 0.0) + v
    for k, a in agg.items():
        n = a["launches"]
        for key in list(a):
            if isinstance(a[key], float):
                a[key] /= n
        a["dram_bytes_per_launch"] = a.get("dram_read_bytes", 0.0) + a.get("dram_write_bytes", 0.0)
        a["note"] = "ncu --set full, mean of %d captured launch(es) in %s" % (n, a["capture"])
        if k in units:
            a["units_per_launch"] = units[k]
            a["dram_bytes_per_unit"] = a["dram_bytes_per_launch"] / units[k]
    prev = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            prev = json.load(f)
    prev.update(agg)
    with open(OUT, "w") as f:
        json.dump(prev, f, indent=1, sort_keys=True)
    print(json.dumps(agg, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
