#!/usr/bin/env python3
"""Exploratory stage-by-stage GPU-vs-oracle comparison (prints, does not assert). Run on a B200 box."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from ft8b200_loader import load
from oracle.pyoracle import Oracle, cand_dtype, msg_dtype, status_dtype, result_dtype
from tools import synth

pkg = load()
O = Oracle()
dev = torch.device("cuda:0")
ctx = pkg.Context(0)
print(pkg.lib().ft8b200_version().decode(), torch.cuda.get_device_name(0))

def t2n(t, dt=None):
    a = t.cpu().numpy()
    return a if dt is None else a.view(dt).reshape(a.shape[:-1])

# ---------- decimator on random bytes (incl. 0x00 / 0xff), 3 streams of different content
rng = np.random.default_rng(11)
nbytes = 2 * 751 * 8 * 40  # 40 super-blocks
iq = rng.integers(0, 256, size=(3, nbytes), dtype=np.uint8)
iq[0, ::53] = 0; iq[1, 7::97] = 255; iq[2] = rng.integers(100, 156, size=nbytes, dtype=np.uint8)
d_iq = torch.from_numpy(iq).to(dev)
d_i, d_q, cnt, peak, y2 = ctx.decimate(d_iq, 3, nbytes, want_y2=True)
torch.cuda.synchronize()
for s in range(3):
    oi, oq, oy2i, oy2q = O.decimate_slot(iq[s], want_y2=True)
    n = int(cnt[s])
    gi, gq = d_i[s].cpu().numpy(), d_q[s].cpu().numpy()
    gy = y2[s].cpu().numpy()
    print(f"decim stream {s}: n={n} oracle_n={oi.size} y2 exact={np.array_equal(gy[:n,0], oy2i) and np.array_equal(gy[:n,1], oy2q)} "
          f"f32 bit-exact I={np.array_equal(gi[:n].view(np.uint32), oi.view(np.uint32))} Q={np.array_equal(gq[:n].view(np.uint32), oq.view(np.uint32))} "
          f"tail zero={not gi[n:].any()} peak ok={float(peak[s]) == max(np.abs(oi).max(), np.abs(oq).max())}")

# ---------- slots: single signal + crowded
slots = []
sig1 = [(O.tones(O.pack_std("CQ", "K1JT", "FN20")), 700.0, 0.5, -10.0)]
I, Q = synth.slot_f32(sig1, 7); I, Q, _ = O.condition(I, Q, 48000); slots.append((I, Q))
I, Q, texts = synth.crowded_band(O, 60, 99); I, Q, _ = O.condition(I, Q, 48000); slots.append((I, Q))
I, Q, texts = synth.crowded_band(O, 25, 5, snr_lo=-18, snr_hi=0); I, Q, _ = O.condition(I, Q, 48000); slots.append((I, Q))
hi = np.stack([s[0] for s in slots]); hq = np.stack([s[1] for s in slots])
d_i = torch.from_numpy(hi).to(dev); d_q = torch.from_numpy(hq).to(dev)
mag = ctx.waterfall(d_i, d_q)
cand, ncand = ctx.find_sync(mag)
ok, stage, status, msg, plain, llr = ctx.decode(mag, cand, ncand, want_plain=True, want_llr=True)
res, nres, umsg, ufreq, uscore = ctx.spots(cand, ncand, ok, msg)
torch.cuda.synchronize()
for s in range(len(slots)):
    o = O.subsystem(hi[s], hq[s])
    gm = mag[s].cpu().numpy()
    nd = int((gm != o["wf"]).sum())
    gc = t2n(cand[s], cand_dtype)[: int(ncand[s])]
    print(f"slot {s}: waterfall diff cells={nd}; ncand gpu={int(ncand[s])} oracle={len(o['cands'])} cand equal={np.array_equal(gc, o['cands'])}")
    bad = 0
    for k, c in enumerate(o["cands"]):
        d = O.decode(o["wf"], c)
        g_ok = int(ok[s, k]); g_st = t2n(status[s], status_dtype)[k]; g_msg = t2n(msg[s], msg_dtype)[k]
        same = (g_ok == d["ok"]) and np.array_equal(plain[s, k].cpu().numpy(), d["plain"]) and \
               np.array_equal(llr[s, k].cpu().numpy().view(np.uint32), d["llr"].view(np.uint32)) and g_st["ldpc_errors"] == d["status"]["ldpc_errors"]
        if d["ok"]:
            same = same and g_msg.tobytes() == d["msg"].tobytes() and g_st.tobytes() == d["status"].tobytes()
        bad += (not same)
    gres = t2n(res[s], result_dtype)
    print(f"   per-candidate mismatches={bad}/{len(o['cands'])}; n_results gpu={int(nres[s])} oracle={o['n']} results equal={gres.tobytes() == o['results'].tobytes()}")
    print("   decoded:", [m["text"].decode() for m in o["msgs"]][:8], "...")

# ---------- raw slot end to end (one real 72 MB slot)
t0 = time.time()
raw = synth.raw_u8([(O.tones(O.pack_std("CQ", "K1JT", "FN20")), 800.0, 0.5, 20.0)], 3)
print("raw synth %.1fs" % (time.time() - t0))
t0 = time.time(); oi, oq = O.decimate_slot(raw); print("oracle decimate %.2fs" % (time.time() - t0), oi.size)
I = np.zeros(48000, np.float32); Q = np.zeros(48000, np.float32); I[:oi.size] = oi; Q[:oq.size] = oq
Ic, Qc, sc = O.condition(I, Q, oi.size)
o = O.subsystem(Ic, Qc)
d_raw = torch.from_numpy(raw).to(dev)
ctx.process_raw(d_raw, 1)
r, n = ctx.fetch_results(1)
print("raw e2e: gpu n=%d oracle n=%d equal=%s" % (n[0], o["n"], r[0].tobytes() == o["results"].tobytes()), r[0][:2])
gi, gq, cnt, peak, _ = ctx.decimate(d_raw, 1, raw.size)
torch.cuda.synchronize()
print("raw decim bit-exact:", np.array_equal(gi[0].cpu().numpy().view(np.uint32), I.view(np.uint32)), np.array_equal(gq[0].cpu().numpy().view(np.uint32), Q.view(np.uint32)), int(cnt[0]))

# ---------- timing (rough)
B = 16
big = d_raw.repeat(B, 1).contiguous()
for _ in range(2): ctx.process_raw(big, B)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); ctx.process_raw(big, B); e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"process_raw x{B}: {ms:.3f} ms -> {B/ms*1e3:.0f} slots/s")
e0.record(); ctx.decimate(big, B, raw.size); e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"decimate x{B}: {ms:.3f} ms -> {B*72.383488e-3/ms:.1f} GB/s algorithmic")
print("launches:", ctx.launches())
