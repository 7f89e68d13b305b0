"""Reporting formats: PSKreporter IPFIX datagram, web-cluster form fields, console table.

CPU tests (the formats are host code; no kernel is involved):
  * oracle restatement == the reference's own postSpots()/webClusterSpots()/printSpots() (oracle/_ref/libref_report.so) on fuzzed records;
  * oracle restatement == the committed golden fixture generated from the reference (tests/golden/report.npz, tools/make_golden.py);
  * the library's host functions (csrc/report.cu) == the oracle, byte for byte, incl. truncation, padding, the 1200-byte cut, wrap-around.
GPU test: records produced by the CUDA path for a batch of slots -> one datagram per slot == the oracle's datagram of the oracle's records.
"""
import numpy as np
import pytest

from conftest import golden
from oracle.pyoracle import ReferenceReport, result_dtype

CALLS = [b"K1JT", b"DL1ABC", b"VK3XYZ/P", b"<...>", b"A", b"PJ4/K1ABC", b"3DA0XYZ", b"YW18FIFA", b"W1AW/QRPP12"[:12], b""]
LOCS = [b"FN20", b"JO62", b"", b"RR73", b"AA00aa"[:6], b"-12", b"R+05", b"73"]


def random_spots(rng, n):
    r = np.zeros(n, result_dtype)
    for k in range(n):
        r[k]["call"] = CALLS[rng.integers(len(CALLS))]
        r[k]["loc"] = LOCS[rng.integers(len(LOCS))]
        r[k]["freq"] = int(rng.integers(-50, 1700)) if rng.random() < 0.9 else int(rng.integers(-2**31, 2**31 - 1))
        r[k]["snr"] = int(rng.integers(0, 120)) if rng.random() < 0.9 else int(rng.integers(-1000, 1000))
    return r


STATIONS = [("W1AW", "FN31", 14074000), ("VE2XYZ/QRP", "FN35ab", 7074000), ("", "", 0), ("N0CALL123456", "AA00", 4294900000)]


def test_oracle_matches_reference_reporting(oracle):
    if not ReferenceReport.available():
        pytest.skip("oracle/_ref/libref_report.so not built (needs /root/reference)")
    ref = ReferenceReport()
    rng = np.random.default_rng(5)
    for trial in range(60):
        n = int(rng.integers(0, ref.max_messages + 1)) if trial else ref.max_messages
        spots = random_spots(rng, n)
        rcall, rloc, dial = STATIONS[trial % len(STATIONS)]
        ut = int(rng.integers(1, 2**32 - 1))
        want = ref.post_spots(spots, rcall, rloc, dial, ut)
        assert len(want) >= 16 + 36 + 60
        rid = int.from_bytes(want[12:16], "big")  # srand(time)/rand() inside the reference: taken from its own packet
        got = oracle.pskreporter_datagram(spots, rcall, rloc, dial, ref.app_version, ut, 1, rid)
        assert got == want, f"trial {trial}: datagram differs"
        assert oracle.print_spots(spots, dial, ut) == ref.print_spots(spots, dial, ut)
        forms = ref.webcluster(spots, rcall, rloc, dial)
        assert len(forms) == n
        for k in range(n):
            assert oracle.webcluster_form(spots[k], rcall, rloc, dial) == forms[k]


def test_oracle_matches_golden_report(oracle):
    g = golden("report")
    spots = g["spots"].view(result_dtype).reshape(-1)
    first = g["first"]
    for c in range(first.size - 1):
        s = spots[first[c]:first[c + 1]]
        rcall, rloc = g["rcall"][c].decode(), g["rloc"][c].decode()
        dial, ut, rid = int(g["dial"][c]), int(g["unixtime"][c]), int(g["random_id"][c])
        want = g["datagrams"][c, :g["datagram_len"][c]].tobytes()
        assert oracle.pskreporter_datagram(s, rcall, rloc, dial, g["app_version"].item().decode(), ut, 1, rid) == want
        assert oracle.print_spots(s, dial, ut) == g["printed"][c].decode()
        for k in range(s.size):
            f = oracle.webcluster_form(s[k], rcall, rloc, dial)
            assert f["_freq"] == g["form_freq"][first[c] + k] and f["_info"] == g["form_info"][first[c] + k]


def test_library_reporting_matches_oracle(pkg, oracle):
    rng = np.random.default_rng(9)
    app = "rtlsdr-ft8d_v0.3.6"
    for trial in range(80):
        n = int(rng.integers(0, 51)) if trial else 50
        spots = random_spots(rng, n)
        rcall, rloc, dial = STATIONS[trial % len(STATIONS)]
        opt = pkg.station(rcall, rloc, dial)
        ut, seq, rid = (int(x) for x in rng.integers(0, 2**32 - 1, 3))
        got, used = pkg.pskreporter_datagram(spots, opt, ut, seq, rid, app)
        want = oracle.pskreporter_datagram(spots, rcall, rloc, dial, app, ut, seq, rid)
        assert got == want, f"trial {trial}"
        assert int.from_bytes(got[2:4], "big") == len(got) and len(got) % 4 == 0 and used <= n
        assert pkg.format_spots(spots, dial, ut) == oracle.print_spots(spots, dial, ut)
        for k in range(min(n, 6)):
            assert pkg.webcluster_form(spots[k], opt) == oracle.webcluster_form(spots[k], rcall, rloc, dial)


def test_library_reporting_edges(pkg, oracle):
    app = "rtlsdr-ft8d_v0.3.6"
    opt = pkg.station("W1AW", "FN31", 14074000)
    # the 1200-byte cut: 50 records of maximum length (34 bytes each) do not all fit; the reference silently drops the rest
    big = np.zeros(50, result_dtype)
    big[:] = (b"PJ4/K1ABC/QR", b"AA00aa", 1234, 30)
    d, used = pkg.pskreporter_datagram(big, opt, 1, 1, 2, app)
    assert used < 50 and d == oracle.pskreporter_datagram(big, "W1AW", "FN31", 14074000, app, 1, 1, 2)
    assert pkg.PSK_MAX_DATAGRAM >= len(d)
    # no spots: header + templates + receiver record + an empty (4-byte) sender set
    d0, used0 = pkg.pskreporter_datagram(big[:0], opt, 7, 1, 2, app)
    assert used0 == 0 and d0 == oracle.pskreporter_datagram(big[:0], "W1AW", "FN31", 14074000, app, 7, 1, 2)
    # a buffer that is too small is an error, never a truncated datagram
    with pytest.raises(pkg.Ft8Error):
        pkg.pskreporter_datagram(big, opt, 1, 1, 2, app, cap=len(d) - 1)
    # records whose char arrays are not terminated (cannot come from the device, but must not overrun)
    raw = np.zeros(1, result_dtype)
    raw.view(np.uint8)[:20] = ord("X")
    d1, _ = pkg.pskreporter_datagram(raw, opt, 1, 1, 2, app)
    assert d1[d1.rindex(b"\x99\x93") + 4] == 13 and b"X" * 13 in d1
    # batch layout: slots without spots send nothing and do not consume a sequence number
    res = np.zeros((4, 50), result_dtype)
    nres = np.array([2, 0, 1, 60], np.int32)   # 60 > max_messages is clamped
    res[0, :2] = big[:2]; res[2, 0] = (b"K1JT", b"FN20", 800, 34); res[3] = big
    grams, k = pkg.pskreporter_batch(res, nres, opt, [10, 11, 12, 13], first_sequence=5, random_id=99, app_version=app)
    assert k == 3 and grams[1] == b""
    assert grams[0] == oracle.pskreporter_datagram(res[0, :2], "W1AW", "FN31", 14074000, app, 10, 5, 99)
    assert grams[2] == oracle.pskreporter_datagram(res[2, :1], "W1AW", "FN31", 14074000, app, 12, 6, 99)
    assert grams[3] == oracle.pskreporter_datagram(res[3], "W1AW", "FN31", 14074000, app, 13, 7, 99)
    assert pkg.format_spots(big[:0], 0, 0) == "No spot 1970-01-01 00:00z\n"


@pytest.mark.gpu
def test_report_from_gpu_records(pkg, ctx, oracle):
    """Whole chain: synthetic slots -> CUDA path -> decoder_results -> datagrams == oracle's datagrams of the oracle's records."""
    import torch
    from tools import synth
    rng = np.random.default_rng(77)
    items, first = [], [0]
    for s in range(6):
        for _ in range(s % 4):   # 0..3 CQ messages per slot (non-CQ messages are not reported by the daemon)
            _, de, _ = synth.random_message(rng)
            items.append((pkg.pack77_std("CQ", de, synth.random_grid(rng)), float(rng.uniform(150.0, 1400.0)), float(rng.uniform(0.2, 0.9)), 0.4))
        first.append(len(items))
    d_i, d_q = ctx.synth_slots(pkg.make_signals(items), first, 1.0, 21)
    peak = torch.maximum(d_i.abs().amax(1), d_q.abs().amax(1))
    ctx.process_conditioned(d_i, d_q, peak)
    res, nres = ctx.fetch_results(6)
    opt = pkg.station("W1AW", "FN31", 14074000)
    times = 1_700_000_000 + 15 * np.arange(6)
    grams, k = pkg.pskreporter_batch(res, nres, opt, times, first_sequence=1, random_id=0xABCD, app_version=None)
    assert k == int((nres > 0).sum()) and k >= 3
    seq = 1
    for s in range(6):
        i_s, q_s, _ = oracle.condition(d_i[s].cpu().numpy(), d_q[s].cpu().numpy(), 48000)
        o = oracle.subsystem(i_s, q_s)
        assert o["n"] == nres[s]
        if o["n"] == 0:
            assert grams[s] == b""
            continue
        want = oracle.pskreporter_datagram(o["results"][:o["n"]], "W1AW", "FN31", 14074000, pkg.lib_app_version(), int(times[s]), seq, 0xABCD)
        assert grams[s] == want
        assert pkg.format_spots(res[s, :nres[s]], 14074000, int(times[s])) == oracle.print_spots(o["results"][:o["n"]], 14074000, int(times[s]))
        seq += 1
