#!/usr/bin/env python3
"""DRAM traffic of the dominant kernel from the ncu summary (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture)
-> the JSON bench.py's roofline.traffic reads.  usage: tools/traffic_from_ncu.py profiles/ncu_summary_r2.json 128 profiles/traffic_r2.json"""
import json, sys
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
summary = json.load(open(sys.argv[1]))
d = summary[next(k for k in sorted(summary) if k.startswith("cic_block_sums_kernel"))]   # "cic_block_sums_kernel<4>": the whole-GPU shape
n = int(sys.argv[2])
tot = d["dram_read"] * UNIT[d["dram_read_unit"]] + d["dram_write"] * UNIT[d["dram_write_unit"]]
alg = 72_000_000 + 47_936 * 8   # raw IQ in, two float rails out, per slot
ms = d["duration"] * {"us": 1e-3, "ms": 1.0, "ns": 1e-6}[d["duration_unit"]]
out = {"kernel": "cic_block_sums_kernel", "slots_per_launch": n, "dram_bytes_per_launch": tot, "dram_bytes_per_slot": tot / n,
       "algorithmic_bytes_per_slot": alg, "ratio": tot / n / alg, "duration_ms_under_ncu": ms,
       "source": "profiles/ncu_raw_cic_block_sums_kernel_r2.csv: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture at %d slots (tools/gpu_r2_prof.sh)" % n}
json.dump(out, open(sys.argv[3], "w"), indent=1)
print(json.dumps(out))
