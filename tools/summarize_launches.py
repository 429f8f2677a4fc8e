"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt
(ncu serialises launches and runs them cold: compare SHARES with bench.py's live CUDA-event times.)"""
import csv
import re
import sys
from collections import defaultdict


SETUP = {"JacToAff", "FixedTableRows", "AffToJac", "k_fq_mul_chain", "k_imad_wide"}   # one-off: CRS tables, peak probes


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        m = re.search(r"k_each<cpg::(\w+)|k_each<\(anonymous namespace\)::(\w+)|<unnamed>::(\w+),", name)
        short = next((g for g in (m.groups() if m else ()) if g), None) or re.sub(r"\(.*", "", name).split("::")[-1]
        val = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit", "ns") in ("us", "usecond"):
            val *= 1e3
        elif r.get("Metric Unit", "ns") in ("ms", "msecond"):
            val *= 1e6
        rows.append((short, val))
    tot = defaultdict(float); cnt = defaultdict(int)
    for k, v in rows:
        tot[k] += v; cnt[k] += 1
    total = sum(tot.values()) or 1.0
    steps = sum(v for k, v in tot.items() if k not in SETUP) or 1.0
    print("# %s: %d launches, %.3f ms in kernels (serialised, cold-cache); %.3f ms outside the one-off setup kernels" % (path, len(rows), total / 1e6, steps / 1e6))
    print("# share = of the step kernels (setup kernels - CRS table construction, peak probes - are listed but not counted)")
    print("%-24s %8s %12s %8s" % ("kernel", "launches", "total_ms", "share"))
    for k in sorted(tot, key=lambda k: -tot[k]):
        share = "  setup" if k in SETUP else "%6.1f%%" % (100.0 * tot[k] / steps)
        print("%-24s %8d %12.3f %8s" % (k, cnt[k], tot[k] / 1e6, share))


if __name__ == "__main__":
    main(sys.argv[1])
