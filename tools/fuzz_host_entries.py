#!/usr/bin/env python3
"""Fuzz of the host-only entry points (no GPU): ft8b200_pack77 / ft8b200_pack77_std on arbitrary byte strings, the report builders on
records whose char fields carry no terminator, with every output capacity from 0 up.  Run it plainly or against the library built
with -fsanitize=address,undefined through FT8B200_LIB_PATH.  usage: tools/fuzz_host_entries.py [seed]"""
import ctypes as C, os, sys
import numpy as np
L = C.CDLL(os.environ.get("FT8B200_LIB_PATH") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rtlsdr-ft8d_b200", "libft8b200.so"))
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
rec = np.dtype([("call", "S13"), ("loc", "S7"), ("freq", "<i4"), ("snr", "<i4")])
assert rec.itemsize == 28
ALPHA = b" ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789+-./?<>"


def text(maxlen, raw=False):
    n = int(rng.integers(0, maxlen + 1))
    if raw:
        return bytes(rng.integers(1, 256, n, dtype=np.uint8))
    return bytes(ALPHA[i] for i in rng.integers(0, len(ALPHA), n))


n = 0
payload = (C.c_uint8 * 10)()
for it in range(4000):
    L.ft8b200_pack77(text(200, raw=it % 3 == 0), payload)
    L.ft8b200_pack77_std(text(40, raw=it % 5 == 0), text(40), text(40), payload)
    k = int(rng.integers(0, 70))
    spots = np.zeros(max(k, 1), rec)
    raw = rng.integers(0, 256, spots.nbytes, dtype=np.uint8)            # char fields without terminators, any freq / snr
    if it % 2:
        raw = np.where(rng.random(raw.size) < 0.3, 0, raw).astype(np.uint8)
    spots = raw.view(rec)
    station = np.frombuffer(bytes(rng.integers(0, 256, 24, dtype=np.uint8)), np.uint8).copy()
    cap = int(rng.integers(0, 1700))
    out = (C.c_uint8 * max(cap, 1))()
    nrep = C.c_uint32(0)
    L.ft8b200_pskreporter_datagram(spots.ctypes.data_as(C.c_void_p), C.c_uint32(k), station.ctypes.data_as(C.c_void_p), None if it % 4 else text(300),
                                   C.c_uint32(int(rng.integers(0, 1 << 32))), C.c_uint32(1), C.c_uint32(7), out, C.c_size_t(cap), C.byref(nrep))
    form = (C.c_uint8 * 138)()
    L.ft8b200_webcluster_form(spots.ctypes.data_as(C.c_void_p), station.ctypes.data_as(C.c_void_p), form)
    cap2 = int(rng.integers(0, 6000))
    txt = C.create_string_buffer(max(cap2, 1))
    L.ft8b200_format_spots(spots.ctypes.data_as(C.c_void_p), C.c_uint32(k), C.c_uint32(int(rng.integers(0, 1 << 32))), C.c_uint32(int(rng.integers(0, 1 << 32))), txt, C.c_size_t(cap2))
    nslots = int(rng.integers(1, 5)); M = int(rng.integers(1, 60))
    batch = rng.integers(0, 256, nslots * M * 28, dtype=np.uint8)
    counts = rng.integers(-3, M + 5, nslots).astype(np.int32)
    stride = int(rng.integers(0, 1700))
    outb = (C.c_uint8 * max(stride * nslots, 1))()
    lens = (C.c_int32 * nslots)()
    times = rng.integers(0, 1 << 32, nslots).astype(np.uint32)
    L.ft8b200_pskreporter_batch(batch.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p), nslots, M, station.ctypes.data_as(C.c_void_p), None,
                                times.ctypes.data_as(C.c_void_p), C.c_uint32(1), C.c_uint32(9), outb, C.c_size_t(stride), lens)
    n += 1
print("fuzzed", n)
