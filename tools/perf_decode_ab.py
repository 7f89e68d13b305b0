#!/usr/bin/env python3
"""A/B of the two belief-propagation kernels (ft8b200_set_decode_variant: 0 = node-centred, 1 = edge-centred) on the two
decode-bound BASELINE configurations, same process, same inputs: whole-call time, the decode stage's own device time, and
that the records are byte-identical.  Writes gpurun_out/perf_decode_ab_<TAG>.json.   usage: tools/perf_decode_ab.py [TAG]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from ft8b200_loader import load
from tools import synth

pkg = load()
TAG = sys.argv[1] if len(sys.argv) > 1 else "x"


def signals(rng, n, f_lo, f_hi, t_lo, t_hi, amp_lo, amp_hi):
    items = []
    for _ in range(n):
        to, de, ex = synth.random_message(rng)
        items.append((pkg.pack77_std(to, de, ex), float(rng.uniform(f_lo, f_hi)), float(rng.uniform(t_lo, t_hi)), float(rng.uniform(amp_lo, amp_hi))))
    return pkg.make_signals(items)


def batch(seed, n_slots, per_slot, *args):
    rng = np.random.default_rng(seed)
    sigs = [signals(rng, per_slot, *args) for _ in range(n_slots)]
    return np.concatenate(sigs), np.concatenate([[0], np.cumsum([s.size for s in sigs])]).astype(np.int32)


def amp_for_snr(snr_db, sigma=1.0):
    return float(np.sqrt(2.0 * sigma * sigma * (2500.0 / 3200.0) * 10.0 ** (snr_db / 10.0)))


def measure(ctx, d_i, d_q, peak, n, reps):
    rows = {}
    keep = None
    for variant in (1, 0, 1, 0):
        pkg.set_decode_variant(variant)
        ctx.set_profiling(False)
        for _ in range(2):
            ctx.process_conditioned(d_i, d_q, peak); ctx.fetch_results(n)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            ctx.process_conditioned(d_i, d_q, peak)
            res, nres = ctx.fetch_results(n)
        torch.cuda.synchronize()
        sec = (time.perf_counter() - t0) / reps
        ctx.set_profiling(True)
        ctx.process_conditioned(d_i, d_q, peak); ctx.fetch_results(n)
        st = ctx.stage_times()
        ctx.set_profiling(False)
        blob = res.tobytes() + nres.tobytes()
        if keep is None:
            keep = blob
        r = rows.setdefault("nodes" if variant == 0 else "edges", {"ms": [], "decode_ms": [], "sync_ms": [], "waterfall_ms": []})
        r["ms"].append(sec * 1e3); r["decode_ms"].append(st["decode"]); r["sync_ms"].append(st["sync"]); r["waterfall_ms"].append(st["waterfall"])
        r["same_records"] = r.get("same_records", True) and blob == keep
        r["spots"] = int(nres.sum())
    pkg.set_decode_variant(0)
    for r in rows.values():
        r["slots_per_s"] = n / (min(r["ms"]) * 1e-3)
    rows["decode_speedup"] = min(rows["edges"]["decode_ms"]) / min(rows["nodes"]["decode_ms"])
    rows["slots"] = n
    return rows


out = {}
ctx = pkg.Context(0)
N = int(os.environ.get("PERF_SLOTS", "4096"))
sig, first = batch(1, N, 1, 100.0, 1400.0, 0.2, 0.8, amp_for_snr(-10.0), amp_for_snr(-10.0))
d_i, d_q = ctx.synth_slots(sig, first, 1.0, 7)
peak = torch.maximum(d_i.abs().amax(1), d_q.abs().amax(1))
out["config1_4_slots_3200sps"] = measure(ctx, d_i, d_q, peak, N, 5)
print(json.dumps(out["config1_4_slots_3200sps"]), flush=True)
del d_i, d_q
ctx.close()

ctx = pkg.Context(0, max_candidates=500, max_messages=200)
N3 = int(os.environ.get("PERF_CROWDED", "1024"))
sig, first = batch(3, N3, 60, 50.0, 1500.0, -0.5, 1.5, amp_for_snr(-24.0), amp_for_snr(5.0))
c_i, c_q = ctx.synth_slots(sig, first, 1.0, 9)
cpeak = torch.maximum(c_i.abs().amax(1), c_q.abs().amax(1))
out["config3_daemon_path_k500"] = measure(ctx, c_i, c_q, cpeak, N3, 3)
print(json.dumps(out["config3_daemon_path_k500"]), flush=True)
ctx.close()

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"perf_decode_ab_{TAG}.json"), "w"), indent=1)
