#!/usr/bin/env python3
"""Fuzz of the on-disk readers (ft8b200_load_wav, ft8b200_load_wav_s16, ft8b200_read_iq_file, ft8b200_read_c2_file; host code, no GPU): random bytes,
RIFF headers with every field out of range, truncated and oversized files.  Run it plainly (a crash is the finding) or with the
library built with -fsanitize=address,undefined through FT8B200_LIB_PATH (profiles/sanitizer_r2.md).  usage: tools/fuzz_file_readers.py [seed]"""
import ctypes as C, os, sys, numpy as np, struct, tempfile
L = C.CDLL(os.environ.get("FT8B200_LIB_PATH", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rtlsdr-ft8d_b200", "libft8b200.so")))
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
d = tempfile.mkdtemp()
hi = (C.c_float * 48000)(); hq = (C.c_float * 48000)(); peak = C.c_float()
sig = (C.c_float * 180000)(); raw = (C.c_int16 * 180000)()
def wav_header(fmt=1, ch=1, rate=12000, bits=16, sub1=16, sub2=1000, align=2):
    return b"RIFF" + struct.pack("<I", (36 + sub2) & 0xFFFFFFFF) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", sub1, fmt, ch, rate, (rate * align) & 0xFFFFFFFF, align, bits) + b"data" + struct.pack("<I", sub2)
n = 0
for it in range(3000):
    kind = it % 6
    if kind == 0: blob = rng.integers(0, 256, int(rng.integers(0, 200)), dtype=np.uint8).tobytes()
    elif kind == 1: blob = wav_header(sub2=int(rng.integers(0, 400000)), align=int(rng.integers(0, 5))) + rng.integers(0, 256, int(rng.integers(0, 3000)), dtype=np.uint8).tobytes()
    elif kind == 2: blob = wav_header(fmt=int(rng.integers(0, 3)), ch=int(rng.integers(0, 3)), bits=int(rng.choice([8, 16, 24])), sub1=int(rng.choice([16, 18, 0])))[: int(rng.integers(0, 50))]
    elif kind == 3: blob = wav_header(sub2=0xFFFFFFF0, align=2) + b"\0" * 100
    elif kind == 4: blob = rng.integers(0, 256, int(rng.integers(0, 40)), dtype=np.uint8).tobytes() + rng.standard_normal(int(rng.integers(0, 200000))).astype(np.float32).tobytes()[: int(rng.integers(0, 800000))]
    else: blob = wav_header(rate=int(rng.integers(0, 1 << 31)), sub2=2 * 179999) + b"\x01\x02" * 179999
    p = os.path.join(d, "f%d" % kind).encode()
    open(p, "wb").write(blob)
    ns = C.c_int(180000); sr = C.c_int(0)
    L.ft8b200_load_wav(sig, C.byref(ns), C.byref(sr), p)
    ns = C.c_int(180000)
    L.ft8b200_load_wav_s16(raw, None, C.byref(ns), C.byref(sr), p)
    L.ft8b200_read_iq_file(p, hi, hq, C.byref(peak))
    df = C.c_double(); ty = C.c_int(); name = C.create_string_buffer(16)
    L.ft8b200_read_c2_file(p, hi, hq, C.byref(peak), C.byref(df), C.byref(ty), name)
    n += 1
print("fuzzed", n)
