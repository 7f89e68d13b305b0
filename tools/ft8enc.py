"""Pure-Python FT8 encoder (message packing for standard messages, CRC-14, LDPC(174,91) parity, Gray/Costas
tone mapping) used ONLY to synthesise benchmark/test inputs without touching oracle/.
Restates ft8_lib/ft8/pack.c:20-232, crc.c:10-63, encode.c:22-125; checked against the oracle and the
reference in tests/test_oracle_vs_ref.py.  The protocol tables are read from the generated header
rtlsdr-ft8d_b200/csrc/ft8_tables.h (tools/gen_tables.py)."""
from __future__ import annotations

import os
import re

import numpy as np

_HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rtlsdr-ft8d_b200", "csrc", "ft8_tables.h")
_tables = {}


def _table(name):
    if not _tables:
        text = open(_HDR).read()
        for m in re.finditer(r"static const uint8_t (\w+)((?:\[\d+\])+) = \{(.*?)\};", text, re.S):
            dims = [int(d) for d in re.findall(r"\[(\d+)\]", m.group(2))]
            vals = [int(v, 0) for v in re.findall(r"0x[0-9a-fA-F]+|\d+", m.group(3))]
            _tables[m.group(1)] = np.array(vals, dtype=np.uint8).reshape(dims)
    return _tables[name]


def crc14(data: bytes, nbits: int) -> int:
    rem, byte = 0, 0
    for b in range(nbits):
        if b % 8 == 0:
            rem ^= data[byte] << 6
            byte += 1
        rem = ((rem << 1) ^ 0x2757) & 0xFFFF if rem & 0x2000 else (rem << 1) & 0xFFFF
    return rem & 0x3FFF


_A1 = " 0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZ"
_A2 = "0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZ"
_A3 = "0123456789"
_A4 = " ABCDEFGHIJKLMNOPQRSTUVWXYZ"


def _pack_call(call: str) -> int:
    if call == "DE":
        return 0
    if call == "QRZ":
        return 1
    if call == "CQ":
        return 2
    n = len(call)
    if n >= 3 and call[2].isdigit() and n <= 6:
        c6 = call.ljust(6)
    elif n >= 2 and call[1].isdigit() and n <= 5:
        c6 = (" " + call).ljust(6)
    else:
        raise ValueError(f"not a standard callsign: {call!r}")
    idx = [_A1.find(c6[0]), _A2.find(c6[1]), _A3.find(c6[2]), _A4.find(c6[3]), _A4.find(c6[4]), _A4.find(c6[5])]
    if min(idx) < 0:
        raise ValueError(f"not a standard callsign: {call!r}")
    v = idx[0]
    for base, i in zip((36, 10, 27, 27, 27), idx[1:]):
        v = v * base + i
    return 2063592 + 4194304 + v


def _pack_extra(x: str) -> int:
    if not x:
        return 32401
    if x in ("RRR", "RR73", "73"):
        return 32400 + {"RRR": 2, "RR73": 3, "73": 4}[x]
    if len(x) == 4 and "A" <= x[0] <= "R" and "A" <= x[1] <= "R" and x[2:].isdigit():
        return ((ord(x[0]) - 65) * 18 + (ord(x[1]) - 65)) * 100 + int(x[2:])
    if x[0] == "R":
        return ((32400 + 35 + int(x[1:])) | 0x8000) & 0xFFFF
    return 32400 + 35 + int(x)


def pack_std(call_to: str, call_de: str, extra: str) -> bytes:
    a, d, g = _pack_call(call_to) << 1, _pack_call(call_de) << 1, _pack_extra(extra)
    b = [a >> 21, a >> 13, a >> 5, (a << 3) | (d >> 26), d >> 18, d >> 10, d >> 2, (d << 6) | (g >> 10), g >> 2, (g << 6) | (1 << 3)]
    return bytes(v & 0xFF for v in b)


def encode174(payload: bytes) -> np.ndarray:
    a91 = bytearray(payload[:10]) + bytearray(2)
    a91[9] &= 0xF8
    crc = crc14(bytes(a91), 82)
    a91[9] |= crc >> 11
    a91[10] = (crc >> 3) & 0xFF
    a91[11] = (crc << 5) & 0xFF
    bits = np.unpackbits(np.frombuffer(bytes(a91), np.uint8))[:91]
    gen = np.unpackbits(_table("kFt8tGen"), axis=1)[:, :91]
    parity = (gen.astype(np.int32) @ bits.astype(np.int32)) & 1
    return np.concatenate([bits, parity.astype(np.uint8)])


def tones(payload: bytes) -> np.ndarray:
    bits = encode174(payload)
    costas, gray = _table("kFt8tCostas"), _table("kFt8tGray")
    out = np.zeros(79, np.uint8)
    k = 0
    for s in range(79):
        if s < 7:
            out[s] = costas[s]
        elif 36 <= s < 43:
            out[s] = costas[s - 36]
        elif s >= 72:
            out[s] = costas[s - 72]
        else:
            out[s] = gray[(bits[k] << 2) | (bits[k + 1] << 1) | bits[k + 2]]
            k += 3
    return out


def tones_ft4(payload: bytes) -> np.ndarray:
    """FT4 channel symbols (105 tones: ramp, 4 Costas groups, 87 two-bit data symbols), ft8_lib/ft8/encode.c:126-195:
    the 77 message bits are scrambled with the protocol's fixed sequence before CRC and parity."""
    xor = _table("kFt4tXor")
    scrambled = bytes(b ^ int(x) for b, x in zip(payload[:10], xor))
    bits = encode174(scrambled)
    costas, gray = _table("kFt4tCostas"), _table("kFt4tGray")
    out = np.zeros(105, np.uint8)
    k = 0
    for s in range(105):
        if s == 0 or s == 104:
            out[s] = 0
        elif s < 5:
            out[s] = costas[0][s - 1]
        elif 34 <= s < 38:
            out[s] = costas[1][s - 34]
        elif 67 <= s < 71:
            out[s] = costas[2][s - 67]
        elif s >= 100:
            out[s] = costas[3][s - 100]
        else:
            out[s] = gray[(bits[k] << 1) | bits[k + 1]]
            k += 2
    return out
