#!/usr/bin/env python3
"""Group an `ncu --metrics gpu__time_duration.sum --csv` launch list by (kernel, grid) and print per-launch device time and
each kernel's share of the step (cold-cache, serialised: compare SHARES with bench.py's live stage times, not absolutes).
usage: tools/launch_summary.py profiles/launches_*.csv"""
import collections, csv, re, sys

def main(path):
    rows = list(csv.DictReader(l for l in open(path, errors="replace") if l.startswith('"')))
    acc = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "").replace("ft8b200::", "")
        a = acc.setdefault((name, r["Grid Size"], r["Block Size"]), [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"]) / 1e3
    groups = collections.defaultdict(float)   # total per distinct batch shape = per second grid dimension (slots)
    print("| kernel | grid | block | launches | us / launch |")
    print("|---|---|---|---|---|")
    for (name, grid, block), (n, us) in acc.items():
        print(f"| `{name}` | {grid} | {block} | {n} | {us / n:.1f} |")

if __name__ == "__main__":
    main(sys.argv[1])
