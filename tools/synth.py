"""Deterministic synthetic FT8 inputs (BASELINE.json configs #1-#5).

Host-side numpy only; used by tests/, bench.py and tools/make_golden.py to feed the CPU oracle and
the CUDA path the SAME bytes.  Nothing here is on the product path.

  slot_f32(...)      one 15 s slot of complex baseband at 3200 sps (what ft8_subsystem() consumes),
                     8-FSK, 512 samples/symbol, 6.25 Hz tone spacing, complex AWGN, SNR in 2500 Hz.
  raw_u8(...)        one 15 s slot of raw RTL-SDR uint8 IQ at 2.4 Msps (what rtlsdr_callback() consumes):
                     the same tone sequence placed at (f - 600 kHz) so that the daemon's fs/4 mixer
                     lands it at f after decimation; offset-128 unsigned bytes with saturation.
"""
from __future__ import annotations

import numpy as np

FS_AUDIO = 3200
SYM_LEN = 512           # samples per symbol at 3200 sps
N_SLOT = 48000
FS_RAW = 2_400_000
RAW_PER_SYM = 384_000   # 0.16 s at 2.4 Msps
RAW_SLOT_SAMPLES = 36_000_000
TONE_HZ = 6.25

_CALL_A = "ABCDEFGHIJKLMNOPQRSTUVWXYZ"


def random_call(rng: np.random.Generator) -> str:
    """A standard callsign the type-1 packer accepts (letter, letter|digit, digit, 1-3 letters)."""
    c = _CALL_A[rng.integers(26)] + (_CALL_A[rng.integers(26)] if rng.random() < 0.6 else "") + str(rng.integers(10))
    c += "".join(_CALL_A[rng.integers(26)] for _ in range(int(rng.integers(1, 4))))
    if len(c) >= 3 and not c[2].isdigit() and not c[1].isdigit():
        c = c[0] + str(rng.integers(10)) + c[2:]
    return c


def random_grid(rng: np.random.Generator) -> str:
    return "ABCDEFGHIJKLMNOPQR"[rng.integers(18)] + "ABCDEFGHIJKLMNOPQR"[rng.integers(18)] + str(rng.integers(10)) + str(rng.integers(10))


def random_message(rng: np.random.Generator):
    """(call_to, call_de, extra) triples covering CQ / grid / report / RR73 forms."""
    kind = rng.integers(5)
    de = random_call(rng)
    if kind <= 1:
        return ("CQ", de, random_grid(rng))
    to = random_call(rng)
    if kind == 2:
        return (to, de, random_grid(rng))
    if kind == 3:
        return (to, de, "%+03d" % int(rng.integers(-24, 10)))
    return (to, de, ["RRR", "RR73", "73"][rng.integers(3)])


def slot_f32(signals, seed: int, noise_sigma: float = 1.0):
    """signals: iterable of (tones[79], f0_hz, t0_sec, snr_db).  Returns (I, Q) float32[48000], NOT yet conditioned.

    SNR is referred to a 2500 Hz bandwidth: amp^2 = 2 sigma^2 (2500/3200) 10^(snr/10), sigma per rail.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    z = (rng.standard_normal(N_SLOT) + 1j * rng.standard_normal(N_SLOT)) * noise_sigma
    t = np.arange(SYM_LEN, dtype=np.float64) / FS_AUDIO
    for tones, f0, t0, snr_db in signals:
        amp = np.sqrt(2.0 * noise_sigma ** 2 * (2500.0 / FS_AUDIO) * 10.0 ** (snr_db / 10.0))
        start = int(round(t0 * FS_AUDIO))
        phase = 0.0
        for k, tone in enumerate(np.asarray(tones, dtype=np.int64)):
            f = f0 + float(tone) * TONE_HZ
            lo = start + k * SYM_LEN
            ph = phase + 2.0 * np.pi * f * t
            phase = (phase + 2.0 * np.pi * f * SYM_LEN / FS_AUDIO) % (2.0 * np.pi)
            a, b = max(lo, 0), min(lo + SYM_LEN, N_SLOT)
            if a < b:
                z[a:b] += amp * np.exp(1j * ph[a - lo:b - lo])
    return z.real.astype(np.float32), z.imag.astype(np.float32)


def raw_u8(signals, seed: int, noise_lsb: float = 30.0, n_samples: int = RAW_SLOT_SAMPLES, dc=(127.5, 127.5)):
    """signals: iterable of (tones[79], f_hz, t0_sec, amp_lsb).  Returns uint8[2*n_samples] interleaved I,Q.

    The signal is synthesised at complex baseband frequency (f - 600 kHz): rtlsdr_callback() multiplies
    sample n by j^n (+fs/4), which moves it to +f in the 3200 sps output.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.empty(2 * n_samples, dtype=np.uint8)
    sig_i = np.zeros(n_samples, dtype=np.float32)
    sig_q = np.zeros(n_samples, dtype=np.float32)
    n = np.arange(RAW_PER_SYM, dtype=np.float64)
    for tones, f_hz, t0, amp in signals:
        start = int(round(t0 * FS_RAW))
        phase = 0.0
        for k, tone in enumerate(np.asarray(tones, dtype=np.int64)):
            f = f_hz + float(tone) * TONE_HZ - 600_000.0
            w = 2.0 * np.pi * f / FS_RAW
            lo = start + k * RAW_PER_SYM
            a, b = max(lo, 0), min(lo + RAW_PER_SYM, n_samples)
            if a < b:
                ph = phase + w * n[a - lo:b - lo]
                sig_i[a:b] += (amp * np.cos(ph)).astype(np.float32)
                sig_q[a:b] += (amp * np.sin(ph)).astype(np.float32)
            phase = (phase + w * RAW_PER_SYM) % (2.0 * np.pi)
    step = 4_000_000
    for o in range(0, n_samples, step):
        e = min(o + step, n_samples)
        ni = rng.standard_normal(e - o, dtype=np.float32) * np.float32(noise_lsb)
        nq = rng.standard_normal(e - o, dtype=np.float32) * np.float32(noise_lsb)
        out[2 * o:2 * e:2] = np.clip(np.rint(sig_i[o:e] + ni + np.float32(dc[0])), 0, 255).astype(np.uint8)
        out[2 * o + 1:2 * e:2] = np.clip(np.rint(sig_q[o:e] + nq + np.float32(dc[1])), 0, 255).astype(np.uint8)
    return out


def crowded_band(oracle, n_signals: int, seed: int, f_lo=50.0, f_hi=1500.0, snr_lo=-24.0, snr_hi=5.0, dt=1.0):
    """BASELINE.json config #3 on the daemon's 0..1600 Hz band: n overlapping messages, random f0/DT/SNR."""
    rng = np.random.Generator(np.random.PCG64(seed ^ 0x5EED))
    sigs, texts = [], []
    for _ in range(n_signals):
        to, de, ex = random_message(rng)
        payload = oracle.pack_std(to, de, ex)
        sigs.append((oracle.tones(payload), float(rng.uniform(f_lo, f_hi)), float(0.5 + rng.uniform(-dt, dt)), float(rng.uniform(snr_lo, snr_hi))))
        texts.append(f"{to} {de} {ex}")
    i_s, q_s = slot_f32(sigs, seed)
    return i_s, q_s, texts


def audio_12k(signals, seed: int, noise_sigma: float = 0.05, n_samples: int = 180_000, fs: int = 12_000, symbol_period: float = 0.16):
    """Real audio at 12 kHz for the ft8_lib monitor path: signals = (tones, f0_hz, t0_sec, amplitude).
    FT8: 79 tones, 0.16 s symbols (1920 samples), 6.25 Hz spacing.  FT4: 105 tones, symbol_period=0.048 (576 samples,
    20.83 Hz spacing), n_samples=90_000.  Phase-continuous FSK, white Gaussian noise.  Returns float32[n_samples]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = rng.standard_normal(n_samples) * noise_sigma
    sym = int(round(fs * symbol_period))
    tone_hz = 1.0 / symbol_period
    t = np.arange(sym, dtype=np.float64) / fs
    for tones, f0, t0, amp in signals:
        start = int(round(t0 * fs))
        phase = 0.0
        for k, tone in enumerate(np.asarray(tones, dtype=np.int64)):
            f = f0 + float(tone) * tone_hz
            lo = start + k * sym
            a, b = max(lo, 0), min(lo + sym, n_samples)
            if a < b:
                x[a:b] += amp * np.cos(phase + 2.0 * np.pi * f * t[a - lo:b - lo])
            phase = (phase + 2.0 * np.pi * f * sym / fs) % (2.0 * np.pi)
    return x.astype(np.float32)


def pack77_fuzz_messages(seed: int, n: int):
    """Message texts for the pack77() parity tests: standard exchanges, special tokens, the 3DA0/3X prefix rewrites, reports
    with and without sign/R, keywords, trailing junk, stray blanks, lower case, free text and empty strings."""
    import random
    rnd = random.Random(seed)
    L, D = "ABCDEFGHIJKLMNOPQRSTUVWXYZ", "0123456789"

    def call():
        k = rnd.random()
        if k < 0.5:
            return rnd.choice(L) + rnd.choice(L + D) + rnd.choice(D) + "".join(rnd.choice(L) for _ in range(rnd.randint(0, 3)))
        if k < 0.7:
            return rnd.choice(L) + rnd.choice(D) + "".join(rnd.choice(L) for _ in range(rnd.randint(1, 3)))
        if k < 0.75:
            return "3DA0" + "".join(rnd.choice(L) for _ in range(rnd.randint(0, 4)))
        if k < 0.8:
            return "3X" + rnd.choice(L + D) + rnd.choice(D) + "".join(rnd.choice(L) for _ in range(rnd.randint(0, 4)))
        if k < 0.9:
            return rnd.choice(["DE", "QRZ", "CQ", "CQ_DX", "CQ 123", "<...>", "PJ4/K1ABC", "K1ABC/P", "W9XYZ/R"])
        return "".join(rnd.choice(L + D + "/ ") for _ in range(rnd.randint(1, 9)))

    def extra():
        k = rnd.random()
        if k < 0.3:
            return rnd.choice(L[:18]) + rnd.choice(L[:18]) + rnd.choice(D) + rnd.choice(D) + rnd.choice(["", "QI", "xx", " 73"])
        if k < 0.6:
            return rnd.choice(["", "R"]) + rnd.choice(["+", "-", ""]) + "".join(rnd.choice(D) for _ in range(rnd.randint(0, 3)))
        if k < 0.8:
            return rnd.choice(["RRR", "RR73", "73", "RRR ", "R", "73 GL", "RR 73"])
        return "".join(rnd.choice(L + D + "+-./? ") for _ in range(rnd.randint(0, 8)))

    msgs = ["", " ", "CQ", "CQ ", "CQ K1JT", "CQ K1JT ", "CQ K1JT FN20", "CQ K1JT FN20QI", "CQ  K1JT FN20", " CQ K1JT FN20", "hello world",
            "TNX BOB 73 GL", "K1", "K1 W2", "A", "AB", "3X", "3DA0", "3DA0 K1ABC FN20", "DE K1ABC -07", "QRZ W9XYZ R-15", "K1ABC W9XYZ R+",
            "K1ABC W9XYZ -", "0123456789ABCDEFGHIJ", "+-./?", "a1bcd k1abc fn20"]
    while len(msgs) < n:
        k = rnd.random()
        if k < 0.75:
            m = call() + " " + call() + (" " + extra() if rnd.random() < 0.8 else "")
        elif k < 0.9:
            m = "".join(rnd.choice(L + D + "+-./? abc") for _ in range(rnd.randint(0, 16)))
        else:
            m = " " * rnd.randint(0, 2) + call() + " " * rnd.randint(1, 2) + call() + " " * rnd.randint(0, 2) + extra()
        msgs.append(m)
    return msgs[:n]


# ---- 77-bit payloads of every type unpack77() knows (and rejects), for fuzzing the device unpacker -------------------
NTOK = 2063592          # unpack.c:12-13
MAX22 = 4194304


def bits_to_payload(fields) -> np.ndarray:
    """fields: [(value, nbits), ...] MSB first, 77 bits in total -> 10 bytes (bits 77..79 zero)."""
    v = 0
    total = 0
    for val, nb in fields:
        assert 0 <= val < (1 << nb), (val, nb)
        v = (v << nb) | int(val)
        total += nb
    assert total == 77, total
    return np.frombuffer((v << 3).to_bytes(10, "big"), np.uint8).copy()


def _n28_of_kind(rng, kind: str) -> int:
    if kind == "token":
        return int(rng.integers(0, 3))                       # DE, QRZ, CQ
    if kind == "cq_nnn":
        return int(rng.integers(3, 1003))                    # CQ 000 .. CQ 999
    if kind == "cq_aaaa":
        return int(rng.choice([1003, 1004, 1030, 1003 + 27 ** 3, 532443, int(rng.integers(1003, 532444))]))
    if kind == "invalid":
        return int(rng.choice([532444, NTOK - 1, int(rng.integers(532444, NTOK))]))    # unpack_callsign returns -1
    if kind == "hashed":
        return int(rng.choice([NTOK, NTOK + MAX22 - 1, int(rng.integers(NTOK, NTOK + MAX22))]))  # <...>
    return int(rng.choice([NTOK + MAX22, (1 << 28) - 1, int(rng.integers(NTOK + MAX22, 1 << 28))]))  # a standard call sign


def payload_fuzz(seed: int, n: int):
    """n payloads cycling through every branch of unpack77 (unpack.c:18-427): free text (i3=0,n3=0), telemetry (n3=5), the
    undefined n3, standard messages i3=1/2 with every token class for both call fields (DE/QRZ/CQ, CQ nnn, CQ aaaa, the
    invalid gap, hashed calls, plain calls), /R and /P flags, grids, blank/RRR/RR73/73, signed reports with and without R,
    non-standard calls (i3=4: every flip/rpt/cq), and i3 = 3, 5, 6, 7 (rejected).  -> (uint8[n, 10], [label, ...])."""
    rng = np.random.default_rng(seed)
    kinds = ["token", "cq_nnn", "cq_aaaa", "invalid", "hashed", "call"]
    out = np.zeros((n, 10), np.uint8)
    labels = []
    for k in range(n):
        sel = k % 16
        if sel == 0:      # free text: 71 random bits (values beyond 42^13 still unpack), n3 = 0
            out[k] = bits_to_payload([(int(rng.integers(0, 1 << 62)) | (int(rng.integers(0, 1 << 9)) << 62), 71), (0, 3), (0, 3)])
            labels.append("free")
        elif sel == 1:    # free text with few characters (leading blanks are trimmed), incl. the empty text
            m = int(rng.integers(0, 42 ** int(rng.integers(0, 5))))
            out[k] = bits_to_payload([(m, 71), (0, 3), (0, 3)])
            labels.append("free_short")
        elif sel == 2:
            out[k] = bits_to_payload([(int(rng.integers(0, 1 << 62)) | (int(rng.integers(0, 1 << 9)) << 62), 71), (5, 3), (0, 3)])
            labels.append("telemetry")
        elif sel == 3:    # i3 = 0 with an n3 the reference does not unpack
            out[k] = bits_to_payload([(int(rng.integers(0, 1 << 62)), 71), (int(rng.choice([1, 2, 3, 4, 6, 7])), 3), (0, 3)])
            labels.append("n3_reject")
        elif sel == 4:
            out[k] = bits_to_payload([(int(rng.integers(0, 1 << 62)), 74), (int(rng.choice([3, 5, 6, 7])), 3)])
            labels.append("i3_reject")
        elif sel in (5, 6):   # non-standard call: 12-bit hash, 58-bit call, flip, rpt, cq
            n58 = int(rng.integers(0, 1 << 58)) if sel == 5 else int(rng.integers(0, 38 ** int(rng.integers(1, 8))))
            out[k] = bits_to_payload([(int(rng.integers(0, 1 << 12)), 12), (n58, 58), (int(rng.integers(0, 2)), 1), (int(rng.integers(0, 4)), 2),
                                      (int(rng.integers(0, 2)), 1), (4, 3)])
            labels.append("nonstd")
        else:             # standard, i3 = 1 or 2
            ka, kb = kinds[int(rng.integers(0, 6))], kinds[int(rng.integers(0, 6))]
            if sel >= 12:     # mostly decodable ones
                ka = kinds[int(rng.choice([0, 1, 2, 4, 5]))]
                kb = kinds[int(rng.choice([4, 5, 5]))]
            g_sel = int(rng.integers(0, 4))
            if g_sel == 0:
                g = int(rng.integers(0, 32401))
            elif g_sel == 1:
                g = 32400 + int(rng.integers(1, 5))          # blank, RRR, RR73, 73
            elif g_sel == 2:
                g = 32400 + 35 + int(rng.integers(-30, 31))  # reports -30 .. +30
            else:
                g = int(rng.integers(32405, 32768))          # every remaining value (reports beyond two digits print as int_to_dd prints them)
            out[k] = bits_to_payload([(_n28_of_kind(rng, ka), 28), (int(rng.integers(0, 2)), 1), (_n28_of_kind(rng, kb), 28), (int(rng.integers(0, 2)), 1),
                                      (int(rng.integers(0, 2)), 1), (g, 15), (int(rng.integers(1, 3)), 3)])
            labels.append("std_%s_%s" % (ka, kb))
    return out, labels
