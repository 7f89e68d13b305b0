#!/usr/bin/env python3
"""Read `ncu --page raw --csv` exports (one kernel launch each) and print the handful of metrics the design cares about.
usage: tools/ncu_summary.py gpurun_out/ncu_raw_*.csv [--json out.json]"""
import csv, json, sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct2"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue_active_elapsed_pct"),
    ("sm__inst_executed.avg.per_cycle_active", "ipc_active"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__instruction_throughput.avg.pct_of_peak_sustained_active", "inst_tp_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("launch__occupancy_limit_warps", "occ_lim_warps"),
    ("sm__maximum_warps_per_active_cycle_pct", "theo_occ_pct"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "smem_dyn"),
    ("launch__shared_mem_per_block_static", "smem_static"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("smsp__pcsamp_warps_issue_stalled_short_scoreboard", "stall_short_sb"),
    ("smsp__pcsamp_warps_issue_stalled_long_scoreboard", "stall_long_sb"),
    ("smsp__pcsamp_warps_issue_stalled_barrier", "stall_barrier"),
    ("smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "stall_math_throttle"),
    ("smsp__pcsamp_warps_issue_stalled_mio_throttle", "stall_mio_throttle"),
    ("smsp__pcsamp_warps_issue_stalled_lg_throttle", "stall_lg_throttle"),
    ("smsp__pcsamp_warps_issue_stalled_wait", "stall_wait"),
    ("smsp__pcsamp_warps_issue_stalled_not_selected", "stall_not_selected"),
    ("smsp__pcsamp_warps_issue_stalled_selected", "stall_selected"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe_fma_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe_lsu_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe_xu_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
]

def read(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        d[h] = (v, u)
    return d

def main():
    out = {}
    argv = sys.argv[1:]
    json_out = None
    if "--json" in argv:
        k = argv.index("--json")
        json_out = argv[k + 1]
        del argv[k:k + 2]
    args = [a for a in argv if not a.startswith("--")]
    for path in args:
        try:
            d = read(path)
        except IndexError:
            print(f"== {path}: no kernel captured (skipped)")
            continue
        name = d.get("Kernel Name", ("?", ""))[0].split("(")[0]
        head, lt, targs = name.partition("<unnamed>::")      # ft8b200::<unnamed>::kernel<args> -> kernel<args>
        name = targs if lt else name.split("::")[-1]
        for tag in ("_12k_", "_ft4_"):   # the same kernel captured at another geometry keeps its own entry
            if tag in path:
                name += tag.rstrip("_")
        rec = {}
        for key, short in KEYS:
            if key in d:
                v, u = d[key]
                try:
                    rec[short] = float(v.replace(",", ""))
                except ValueError:
                    rec[short] = v
                rec[short + "_unit"] = u
        out[name] = rec
        print(f"== {name}  ({path})")
        for k, v in rec.items():
            if not k.endswith("_unit"):
                print(f"   {k:22s} {v} {rec.get(k + '_unit', '')}")
    if json_out:
        json.dump(out, open(json_out, "w"), indent=1)

if __name__ == "__main__":
    main()
