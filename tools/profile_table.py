#!/usr/bin/env python3
"""Markdown table of the per-kernel ncu summary (tools/ncu_summary.py --json) for profiles/summary_*.md.
usage: tools/profile_table.py profiles/ncu_summary_r2.json"""
import json, sys
d = json.load(open(sys.argv[1]))
print("| kernel | duration | warp instr. | IPC (active) | issue active % | warps active % | regs | DRAM % | smem bank conflicts |")
print("|---|---|---|---|---|---|---|---|---|")
for name in sorted(d):
    k = d[name]
    dur = k["duration"] * {"us": 1.0, "ms": 1e3, "ns": 1e-3}[k["duration_unit"]]
    dur_s = "%.3f ms" % (dur / 1e3) if dur >= 1000 else "%.1f µs" % dur
    print("| `%s` | %s | %.3g | %.2f | %.1f | %.1f | %d | %.1f | %.3g |" % (name, dur_s, k["warp_insts"], k["ipc_active"], k["issue_active_pct"], k["warps_active_pct"],
                                                                    k["regs"], k["dram_pct"], k.get("smem_bank_conflicts", 0.0)))
