#!/usr/bin/env python3
"""Per-source-line hot spots from an .ncu-rep captured with --import-source on (needs ncu on PATH).
usage: tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv, io, subprocess, sys

def num(x):
    try:
        return int(x)
    except ValueError:
        return 0

def main(rep, top=30):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    h = rows[hi]
    si, ie = h.index("# Samples"), h.index("Instructions Executed")
    stall = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    lines, cur_file = [], None
    tot_s = tot_i = 0
    agg = {c: 0 for _, c in stall}
    for r in rows:
        if r and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if len(r) <= si or r[0] == "Line No":
            continue
        if r[0] != "":   # source line row (aggregated over its SASS)
            lines.append((num(r[si]), num(r[ie]), cur_file, r[0], r[1].strip(), [(c, num(r[i])) for i, c in stall]))
        elif r[2] not in ("...", ""):
            tot_s += num(r[si]); tot_i += num(r[ie])
            for i, c in stall:
                agg[c] += num(r[i])
    print(f"total samples {tot_s}, warp instructions {tot_i}")
    print("stalls:", ", ".join(f"{c[6:]}={v * 100 // max(tot_s, 1)}%" for c, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    for s, n, f, ln, src, st in sorted(lines, key=lambda x: -x[0])[:top]:
        st = sorted(st, key=lambda x: -x[1])[:2]
        print(f"{s * 100.0 / max(tot_s, 1):5.1f}% inst {n * 100.0 / max(tot_i, 1):5.1f}%  {f}:{ln:>4}  {src[:100]}   [{', '.join(f'{c[6:]}={v}' for c, v in st if v)}]")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
