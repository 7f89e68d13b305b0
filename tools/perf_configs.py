#!/usr/bin/env python3
"""Throughput of the BASELINE.json configurations that are NOT the bench line (they are parity-test cases; these numbers are
secondary evidence, written to gpurun_out/perf_configs_<TAG>.json):

  #1/#4  batch of independent 3200 sps complex slots (one message each) -> ft8b200_process_conditioned      [slots/s]
  #3 A   crowded band, daemon path: 60 signals per slot, K=500 candidates / 200 messages                      [slots/s]
  #3 B   crowded band, 12 kHz monitor path (200-3000 Hz): ft8b200_decode_audio (decode_ft8 main() batched)    [slots/s]
  #5     receiver streams x 8 consecutive slots through ft8b200_process_raw_streams                           [slots/s, MS/s]

Every number is device work timed between synchronisations (inputs generated on the device, records read back to the host);
the CPU column is the unmodified reference (or the restatement when oracle/_ref is absent) on a few of the same inputs, 1 thread.
usage: tools/perf_configs.py [TAG]
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from ft8b200_loader import load
from tools import synth

pkg = load()
dev = torch.device("cuda:0")
TAG = sys.argv[1] if len(sys.argv) > 1 else "x"
out = {}


def signals(rng, n, f_lo, f_hi, t_lo, t_hi, amp_lo, amp_hi):
    items = []
    for _ in range(n):
        to, de, ex = synth.random_message(rng)
        items.append((pkg.pack77_std(to, de, ex), float(rng.uniform(f_lo, f_hi)), float(rng.uniform(t_lo, t_hi)), float(rng.uniform(amp_lo, amp_hi))))
    return pkg.make_signals(items)


def batch_signals(seed, n_slots, per_slot, *args):
    rng = np.random.default_rng(seed)
    sigs = [signals(rng, per_slot, *args) for _ in range(n_slots)]
    first = np.concatenate([[0], np.cumsum([s.size for s in sigs])]).astype(np.int32)
    return np.concatenate(sigs), first, sigs


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, r


def amp_for_snr(snr_db, sigma):
    # complex noise of variance 2 sigma^2 over 3200 Hz; SNR quoted in 2500 Hz
    return float(np.sqrt(2.0 * sigma * sigma * (2500.0 / 3200.0) * 10.0 ** (snr_db / 10.0)))


from oracle.pyoracle import Oracle
orc = Oracle()

# ---- #1 / #4: independent 3200 sps slots, one message each at -10 dB
ctx = pkg.Context(0)
N = int(os.environ.get("PERF_SLOTS", "4096"))
a = amp_for_snr(-10.0, 1.0)
sig, first, per = batch_signals(1, N, 1, 100.0, 1400.0, 0.2, 0.8, a, a)
d_i, d_q = ctx.synth_slots(sig, first, 1.0, 7)
peak = torch.maximum(d_i.abs().amax(1), d_q.abs().amax(1))
def run14():
    ctx.process_conditioned(d_i, d_q, peak)
    return ctx.fetch_results(N)
sec, (res, nres) = timed(run14)
ncpu = 24
t0 = time.perf_counter()
cpu_n = []
for s in range(ncpu):
    i_s, q_s, _ = orc.condition(d_i[s].cpu().numpy(), d_q[s].cpu().numpy(), 48000)
    cpu_n.append(int(orc.subsystem(i_s, q_s)["n"]))
cpu_sec = (time.perf_counter() - t0) / ncpu
out["config1_4_slots_3200sps"] = {"slots": N, "ms": sec * 1e3, "slots_per_s": N / sec, "decoded_slots": int((nres > 0).sum()),
                                  "cpu_slots_per_s_1thread": 1.0 / cpu_sec, "cpu_same_counts": cpu_n == [int(x) for x in nres[:ncpu]]}
print("config #1/#4:", out["config1_4_slots_3200sps"], flush=True)
ctx.close()

# ---- #3 A: crowded band on the daemon path, K = 500, 200 messages
ctx = pkg.Context(0, max_candidates=500, max_messages=200)
N3 = int(os.environ.get("PERF_CROWDED", "1024"))
sig, first, per = batch_signals(3, N3, 60, 50.0, 1500.0, -0.5, 1.5, amp_for_snr(-24.0, 1.0), amp_for_snr(5.0, 1.0))
c_i, c_q = ctx.synth_slots(sig, first, 1.0, 9)
cpeak = torch.maximum(c_i.abs().amax(1), c_q.abs().amax(1))
def run3a():
    ctx.process_conditioned(c_i, c_q, cpeak)
    return ctx.fetch_results(N3)
sec, (res, nres) = timed(run3a, reps=3)
t0 = time.perf_counter()
cpu_n = []
for s in range(4):
    i_s, q_s, _ = orc.condition(c_i[s].cpu().numpy(), c_q[s].cpu().numpy(), 48000)
    cpu_n.append(int(orc.subsystem(i_s, q_s, max_cand=500, max_msgs=200)["n"]))
cpu_sec = (time.perf_counter() - t0) / 4
out["config3_daemon_path_k500"] = {"slots": N3, "signals_per_slot": 60, "ms": sec * 1e3, "slots_per_s": N3 / sec, "mean_spots_per_slot": float(nres.mean()),
                                   "cpu_slots_per_s_1thread": 1.0 / cpu_sec, "cpu_same_counts": cpu_n == [int(x) for x in nres[:4]]}
print("config #3 (A):", out["config3_daemon_path_k500"], flush=True)
del c_i, c_q
ctx.close()

# ---- #3 B: crowded band on the 12 kHz monitor path (decode_ft8's main(): K = 120, 50 messages, 200-3000 Hz)
ctx = pkg.Context(0)
NB = int(os.environ.get("PERF_AUDIO", "512"))
sig, first, per = batch_signals(4, NB, 60, 200.0, 3000.0, 0.0, 1.5, 0.02, 0.5)
aud = ctx.synth_audio(sig, first, 1, 0.05, 13)
sec, lines = timed(lambda: pkg.decode_audio(ctx, aud, 12000, 1), reps=3)
t0 = time.perf_counter()
same = True
for s in range(3):
    want = orc.decode_ft8_lines(aud[s].cpu().numpy(), 12000, protocol=1)
    same &= [pkg.format_decoded(r) for r in lines[s]] == want
cpu_sec = (time.perf_counter() - t0) / 3
out["config3_monitor_path_12k"] = {"slots": NB, "signals_per_slot": 60, "ms": sec * 1e3, "slots_per_s": NB / sec,
                                   "mean_decodes_per_slot": float(np.mean([len(l) for l in lines])),
                                   "cpu_slots_per_s_1thread": 1.0 / cpu_sec, "cpu_same_lines": bool(same)}
print("config #3 (B):", out["config3_monitor_path_12k"], flush=True)
del aud
ctx.close()
torch.cuda.empty_cache()

# ---- #5: receiver streams x 8 consecutive slots (per GPU: 256/8 = 32 streams x 8 x 72 MB = 18.4 GB)
ctx = pkg.Context(0)
S, K = int(os.environ.get("PERF_STREAMS", "32")), 8
import bench
raw, _ = bench.gen_batch(S * K, 0, dev, ctx)          # [S*K, 72e6]: stream s = rows s*K .. s*K+K-1, contiguous
def run5():
    ctx.process_raw_streams(raw, S, K)
    return ctx.fetch_results(S * K)
sec, (res, nres) = timed(run5, reps=3)
out["config5_streams"] = {"streams": S, "slots_per_stream": K, "input_gb": S * K * 72e6 / 1e9, "ms": sec * 1e3, "slots_per_s": S * K / sec,
                          "msps": S * K * 36.0 / sec, "hbm_gbs_algorithmic": S * K * 72_383_488 / sec / 1e9, "decoded_slots": int((nres > 0).sum())}
print("config #5:", out["config5_streams"], flush=True)
ctx.close()

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"perf_configs_{TAG}.json"), "w"), indent=1)
