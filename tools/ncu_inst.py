#!/usr/bin/env python3
"""ncu_summary JSON -> warp instructions per slot of the issue-bound stages (what bench.py's roofline_extra divides by live time).
usage: tools/ncu_inst.py ncu_summary.json SLOTS_PER_LAUNCH out.json"""
import json, sys
d = json.load(open(sys.argv[1])); n = int(sys.argv[2])
def inst(*names):
    tot = 0.0
    for k, v in d.items():
        if any(k == nm or k.startswith(nm + "<") for nm in names) and "warp_insts" in v:
            tot += v["warp_insts"]
    return tot
out = {"slots_per_launch": n, "source": sys.argv[1],
       "sync": {"warp_inst_per_slot": (inst("sync_score_ft8_kernel") + inst("sync_select_kernel")) / n},
       "decode": {"warp_inst_per_slot": inst("decode_kernel") / n},
       "spots": {"warp_inst_per_slot": inst("spots_kernel") / n},
       "waterfall": {"warp_inst_per_slot": inst("waterfall1024_kernel") / n},
       "monitor": {"warp_inst_per_slot": sum(v.get("warp_insts", 0.0) for k, v in d.items() if k.startswith("monitor_frames_kernel") and k.endswith("_12k")) / n},
       "sync960": {"warp_inst_per_slot": (d.get("sync_score_ft8_kernel_12k", {}).get("warp_insts", 0.0) + d.get("sync_select_kernel_12k", {}).get("warp_insts", 0.0)) / n}}
json.dump(out, open(sys.argv[3], "w"), indent=1)
print(json.dumps(out))
