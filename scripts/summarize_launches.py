#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.

  python scripts/summarize_launches.py gpurun_out/launches.csv ["header line" ...] > profiles/ncu_launches_summary.txt
"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    iname, igrid, ival, iunit, imetric = (hdr.index(k) for k in ("Kernel Name", "Grid Size", "Metric Value", "Metric Unit", "Metric Name"))
    tot = defaultdict(lambda: [0, 0.0])
    for r in rd:
        if len(r) <= ival or r[imetric] != "gpu__time_duration.sum":
            continue
        v = float(r[ival].replace(",", ""))
        unit = r[iunit]
        us = v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(unit, 1e-3)
        name = r[iname].split("(")[0]
        key = (name, r[igrid])
        tot[key][0] += 1
        tot[key][1] += us
    total = sum(v[1] for v in tot.values())
    n = sum(v[0] for v in tot.values())
    for h in sys.argv[2:]:
        print("# " + h)
    print(f"# total {total / 1e3:.2f} ms over {n} launches")
    for (name, grid), (cnt, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        if us / total < 0.0015:
            continue
        print(f"{us / 1e3:10.2f} ms {100 * us / total:5.1f}%  n={cnt:5d} avg={us / cnt:9.1f} us  {name} grid={grid}")


if __name__ == "__main__":
    main()
