#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libref_*.so, built from
/root/reference by `make -C oracle ref`).  Run in the build container only; the fixtures are committed
so the GPU box (which has no /root/reference) can check the oracle and the CUDA path against them.

Every fixture holds the INPUT bytes/samples and the reference's outputs at each stage tap:
  decim_*      uint8 IQ -> rtlsdr_callback() float outputs (65536-byte calls, zero initial state)
  slot_*       conditioned 3200 sps slot -> ft8_subsystem(): waterfall, candidate list, per-candidate
               bp_decode input LLRs / hard decisions / status / message, decoder_results[], n_results
  kat          pack77/ft8_encode known answers (rtlsdr_ft8d.c:919-923), CRC, window table
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle.pyoracle import Oracle, Reference
from tools import synth

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
O = Oracle()


def slot_fixture(name, i_s, q_s, variant, store_input=True):
    R = Reference(variant)
    r = R.subsystem(i_s, q_s)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), i=i_s if store_input else np.zeros(0, np.float32),
                        q=q_s if store_input else np.zeros(0, np.float32), kmax=R.kmax, mmax=R.mmax, n=r["n"], results=r["results"], wf=r["wf"],
                        cands=r["cands"], dec_ok=r["dec_ok"], dec_status=r["dec_status"], dec_msg=r["dec_msg"], llr=r["llr"], plain=r["plain"],
                        bp_errors=r["bp_errors"])
    print(name, "cands", len(r["cands"]), "n_results", r["n"], [m["text"].decode() for m, ok in zip(r["dec_msg"], r["dec_ok"]) if ok][:6])


def report_fixture():
    """Reporting formats from the reference's own postSpots()/webClusterSpots()/printSpots() (oracle/_ref/libref_report.so)."""
    from oracle.pyoracle import ReferenceReport, result_dtype
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_report as T
    ref = ReferenceReport()
    rng = np.random.default_rng(2024)
    spots, first, grams, lens, printed, ff, fi, meta = [], [0], np.zeros((8, 2048), np.uint8), [], [], [], [], []
    for c in range(8):
        n = [0, 1, 3, 7, 20, 50, 50, 12][c]
        s = T.random_spots(rng, n)
        if c == 6:
            s[:] = (b"PJ4/K1ABC/QR", b"AA00aa", 1234, 30)   # maximum-length records: the 1200-byte cut
        rcall, rloc, dial = T.STATIONS[c % len(T.STATIONS)]
        ut = int(rng.integers(1, 2**32 - 1))
        d = ref.post_spots(s, rcall, rloc, dial, ut)
        grams[c, :len(d)] = np.frombuffer(d, np.uint8)
        lens.append(len(d))
        printed.append(ref.print_spots(s, dial, ut).encode())
        for f in ref.webcluster(s, rcall, rloc, dial):
            ff.append(f["_freq"]); fi.append(f["_info"])
        meta.append((rcall.encode(), rloc.encode(), dial, ut, int.from_bytes(d[12:16], "big")))
        spots.append(s); first.append(first[-1] + n)
    np.savez_compressed(os.path.join(OUT, "report.npz"), spots=np.concatenate(spots).view(np.uint8), first=np.array(first, np.int32),
                        rcall=np.array([m[0] for m in meta]), rloc=np.array([m[1] for m in meta]), dial=np.array([m[2] for m in meta], np.uint32),
                        unixtime=np.array([m[3] for m in meta], np.uint32), random_id=np.array([m[4] for m in meta], np.uint32),
                        datagrams=grams, datagram_len=np.array(lens, np.int32), printed=np.array(printed),
                        form_freq=np.array(ff), form_info=np.array(fi), app_version=np.array(ref.app_version.encode()))
    print("report.npz:", lens)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "report":
        return report_fixture()
    # --- decimator: random bytes with forced 0x00 / 0xff (int8 wrap quirk), 10 super-blocks + ragged tail
    rng = np.random.default_rng(2024)
    nbytes = 12016 * 10 + 8 * 100
    iq = rng.integers(0, 256, size=nbytes, dtype=np.uint8)
    iq[::41] = 0
    iq[5::67] = 255
    R = Reference("k120", fresh=True)
    for o in range(0, nbytes, 65536):
        R.callback(iq[o:o + 65536])
    ri, rq, n = R.rx()
    np.savez_compressed(os.path.join(OUT, "decim_random.npz"), iq=iq, i=ri[:n], q=rq[:n])
    print("decim_random", n)

    # --- config #1: one message at -10 dB
    sig = [(O.tones(O.pack_std("CQ", "K1JT", "FN20")), 700.0, 0.5, -10.0)]
    i_s, q_s = synth.slot_f32(sig, 7)
    i_s, q_s, _ = O.condition(i_s, q_s, 48000)
    slot_fixture("slot_single", i_s, q_s, "k120")
    # --- config #3: crowded band, 60 signals, K=500 / M=200
    i_s, q_s, texts = synth.crowded_band(O, 60, 99)
    i_s, q_s, _ = O.condition(i_s, q_s, 48000)
    slot_fixture("slot_crowded_k500", i_s, q_s, "k500")
    slot_fixture("slot_crowded_k120", i_s, q_s, "k120", store_input=False)  # same input as slot_crowded_k500

    # --- real-world 12 kHz recordings (the reference's own test WAVs) through the reference's own main():
    #     input PCM + the exact stdout lines of `decode_ft8 file.wav` (ft8_lib/decode_ft8.c:226-409)
    from oracle.pyoracle import ReferenceMonitor
    import wave
    mon = ReferenceMonitor()
    pcm, lines, names = [], [], []
    for rel in ("ft8_lib/tests/191111_110130.wav", "ft8_lib/tests/20m_busy/test_06.wav", "ft8_lib/tests/websdr_test4.wav"):
        path = os.path.join("/root/reference", rel)
        with wave.open(path, "rb") as w:
            assert (w.getnchannels(), w.getsampwidth(), w.getframerate()) == (1, 2, 12000)
            pcm.append(np.frombuffer(w.readframes(w.getnframes()), np.int16).copy())
        out = mon.decode_ft8_stdout(path)
        lines.append("\n".join(out))
        names.append(rel)
        print(rel, len(out), "messages")
    np.savez_compressed(os.path.join(OUT, "recordings_12k.npz"), names=np.array(names), pcm=np.stack(pcm), lines=np.array(lines))

    # --- known answers
    R = Reference("k120")
    p = R.pack77("CQ K1JT FN20QI")
    np.savez_compressed(os.path.join(OUT, "kat.npz"), packed=np.frombuffer(p, np.uint8), tones=R.tones(p), window=R.window(),
                        crc_test3=np.array([R.crc(bytes([0x11, 0, 0, 0, 0, 0x0E, 0x10, 0x04, 0x01, 0x00, 0, 0]), 76)]))
    assert p.hex() == "000000204dfcdc8a1408"  # rtlsdr_ft8d.c:921
    assert "".join(map(str, R.tones(p))) == "3140652000000001005477547106035036373140652547441342116056460065174427143140652"  # :922
    report_fixture()
    # pack77(): message texts -> the reference's payloads (ft8_lib/ft8/pack.c:284-301)
    msgs = synth.pack77_fuzz_messages(11, 400)
    np.savez_compressed(os.path.join(OUT, "pack77.npz"), msgs=np.array(msgs), packed=np.stack([np.frombuffer(R.pack77(m), np.uint8) for m in msgs]))


if __name__ == "__main__":
    main()
