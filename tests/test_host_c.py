"""The C host side (host/ft8d_host.c): the reference's language over libft8b200.so's C ABI.

Without a GPU: it compiles as strict C (gnu17, -Wall -Wextra -Werror) against include/ft8b200.h alone, links to the
library, includes no CUDA header, and stops loudly when there is no device.
With a GPU (`-m gpu`): its stdout -- the daemon's self-test, file decode and live-receive flows and ft8_lib's decode_ft8 --
equals what the CPU oracle / the unmodified reference's golden stdout say for the same inputs."""
import os
import subprocess
import wave

import numpy as np
import pytest

from conftest import golden
from tools import ft8enc, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_DIR = os.path.join(ROOT, "host")
HOST = os.path.join(HOST_DIR, "ft8d_host")
T0 = 1700000000  # fixed clock for the "No spot <UTC>" line


@pytest.fixture(scope="module")
def host(pkg):
    pkg.lib()  # builds libft8b200.so if it is missing
    subprocess.check_call(["make", "-C", HOST_DIR, "-s"])
    return HOST


def run(host, *args, check=True):
    p = subprocess.run([host, *map(str, args)], capture_output=True, text=True, timeout=600)
    if check:
        assert p.returncode == 0, p.stderr
    return p


# ------------------------------------------------------------------------------------------------- no GPU needed
def test_host_is_strict_c_over_the_abi_only(host):
    src = open(os.path.join(HOST_DIR, "ft8d_host.c")).read()
    includes = [l.split()[1] for l in src.splitlines() if l.startswith("#include")]
    assert '"ft8b200.h"' in includes
    assert not [i for i in includes if "cuda" in i.lower() or "oracle" in i.lower()], "the host sees the C ABI only"
    # from scratch with the strict flags, as a C (not C++) translation unit
    out = subprocess.run(["gcc", "-std=gnu17", "-Wall", "-Wextra", "-Werror", "-pedantic-errors", "-fsyntax-only", "-x", "c",
                          "-I", os.path.join(ROOT, "include"), os.path.join(HOST_DIR, "ft8d_host.c")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    needed = subprocess.run(["ldd", host], capture_output=True, text=True).stdout
    assert "libft8b200.so" in needed and "not found" not in needed


def test_host_usage_and_loud_failure_without_a_gpu(host, tmp_path):
    assert run(host, check=False).returncode == 1
    assert run(host, "receive", "-b", "12", "x.u8", check=False).returncode == 1  # callback sizes are multiples of 8 bytes
    assert run(host, "wav", tmp_path / "missing.wav", check=False).returncode == 1
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the loud-failure path cannot be shown")
    for args in (("selftest",), ("decode", os.path.join(ROOT, "selftest.iq")), ("batch", 1, 1, 1)):
        p = run(host, *args, check=False)
        assert p.returncode != 0 and "no CPU fallback" in p.stderr and "SUCCESS" not in p.stdout


# ------------------------------------------------------------------------------------------------- on the B200
@pytest.mark.gpu
def test_selftest_flow(host, oracle, tmp_path):
    """`rtlsdr_ft8d -t` (rtlsdr_ft8d.c:913-972,1181-1190): KAT message -> tones (both checked inside the program against the
    reference's known answers) -> FSK -> initFFTW/ft8_subsystem -> table; the table equals the oracle's on the saved samples."""
    iq = tmp_path / "selftest_host.iq"
    p = run(host, "-f", 0, "-T", T0, "selftest", iq)
    d = np.fromfile(iq, np.float32)
    assert d.size == 96000
    o = oracle.subsystem(np.ascontiguousarray(d[0::2]), np.ascontiguousarray(-d[1::2]))
    assert o["n"] >= 1 and o["results"][0]["call"] == b"K1JT" and o["results"][0]["loc"] == b"FN20"
    assert p.stdout == oracle.print_spots(o["results"][: o["n"]], 0, T0) + "Self-test SUCCESS!\n"


@pytest.mark.gpu
def test_decode_recorded_files_flow(host, oracle, tmp_path):
    """`rtlsdr_ft8d -r file` (decodeRecordedFile, :859-887) for several files in one batch, incl. the reference's own selftest.iq."""
    sigs = [(ft8enc.tones(ft8enc.pack_std("CQ", "K1JT", "FN20")), 700.0, 0.5, -8.0),
            (ft8enc.tones(ft8enc.pack_std("CQ", "DL1ABC", "JO62")), 1210.0, 0.8, -5.0),
            (ft8enc.tones(ft8enc.pack_std("K1ABC", "W9XYZ", "-15")), 410.0, 0.3, -3.0)]
    i_s, q_s = synth.slot_f32(sigs, 21)
    inter = np.empty(96000, np.float32)
    inter[0::2] = i_s * np.float32(3.0); inter[1::2] = -(q_s * np.float32(3.0))
    inter.tofile(tmp_path / "a.iq")
    with open(tmp_path / "b.c2", "wb") as f:
        f.write(b"000000_0000.c2".ljust(14, b"\0") + np.int32(2).tobytes() + np.float64(7.074).tobytes() + inter[:2 * 45000].tobytes())
    paths = [tmp_path / "a.iq", tmp_path / "b.c2", os.path.join(ROOT, "selftest.iq"), tmp_path / "missing.iq"]
    want = ""
    for path, n in zip(paths, (48000, 45000, 48000, 0)):
        want += f"Number of samples: {n}\n"
        if n == 0:
            continue
        d = np.fromfile(path, np.float32, offset=26 if str(path).endswith(".c2") else 0)
        fi = np.zeros(48000, np.float32); fq = np.zeros(48000, np.float32)
        fi[:n] = d[0:2 * n:2]; fq[:n] = -d[1:2 * n:2]
        ci, cq, _ = oracle.condition(fi, fq, n)  # readRawIQfile's normalisation, :762-778
        o = oracle.subsystem(ci, cq)
        want += oracle.print_spots(o["results"][: o["n"]], 7074000, T0)
    p = run(host, "-f", 7074000, "-T", T0, "decode", *paths)
    assert p.stdout == want
    assert "K1JT" in p.stdout and "DL1ABC" in p.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("own_stream", [False, True])
def test_live_receive_flow(host, oracle, tmp_path, own_stream):
    """rtlsdr_callback(buf, 65536, NULL) the way librtlsdr drives it, the 15 s flip and decoder() (:76-285,1336-1354) on a
    recording of 1 slot + 0.2 slot: the first slot decodes to the oracle's table (filter run sample by sample on the CPU),
    the short remainder is skipped like the reference skips a partial first buffer."""
    first = synth.raw_u8([(ft8enc.tones(ft8enc.pack_std("CQ", "K1JT", "FN20")), 800.0, 0.5, 20.0)], 3)
    rng = np.random.default_rng(5)
    tail = rng.integers(96, 160, size=65536 * 220, dtype=np.uint8)
    path = tmp_path / "rx.u8"
    with open(path, "wb") as f:
        f.write(first.tobytes()); f.write(tail.tobytes())
    # the flip happens after the callback that carries the stream past 72 000 000 bytes: those bytes belong to slot 0
    n_first = -(-first.size // 65536) * 65536
    stream = np.concatenate([first, tail])
    st = oracle.new_decim()
    oi, oq = [], []
    for o in range(0, n_first, 65536):
        a, b = oracle.decim_feed(st, stream[o:o + 65536], 64)
        oi.append(a); oq.append(b)
    oi = np.concatenate(oi); oq = np.concatenate(oq)
    fi = np.zeros(48000, np.float32); fq = np.zeros(48000, np.float32)
    n = min(oi.size, 48000)
    fi[:n] = oi[:n]; fq[:n] = oq[:n]
    ci, cq, _ = oracle.condition(fi, fq, n)
    o = oracle.subsystem(ci, cq)
    assert o["n"] >= 1
    rest = (stream.size - n_first) // 1502
    want = f"slot 0: {n} samples\n" + oracle.print_spots(o["results"][: o["n"]], 14074000, T0)
    args = ["-T", T0, "receive"] + (["-s"] if own_stream else []) + [path]
    p = run(host, *args)
    lines = p.stdout.splitlines(keepends=True)
    assert "".join(lines[:-1]) == want
    assert lines[-1].startswith("slot 1: ") and lines[-1].endswith("signal too short, skipped\n")
    assert abs(int(lines[-1].split()[2]) - rest) <= 1
    assert "K1JT" in p.stdout


@pytest.mark.gpu
def test_decode_ft8_flow_on_real_recordings(host, tmp_path):
    """ft8_lib's decode_ft8 main() (decode_ft8.c:226-409) written against the drop-in monitor_* / ft8_find_sync / ft8_decode:
    same stdout as the unmodified reference on three of its real-world recordings (tests/golden/recordings_12k.npz)."""
    g = golden("recordings_12k")
    for k, pcm in enumerate(g["pcm"]):
        path = tmp_path / f"rec{k}.wav"
        with wave.open(str(path), "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(12000)
            w.writeframes(np.ascontiguousarray(pcm, np.int16).tobytes())
        p = run(host, "wav", path)
        assert p.stdout.splitlines() == str(g["lines"][k]).split("\n"), str(g["names"][k])


@pytest.mark.gpu
def test_executor_from_c(host):
    p = run(host, "batch", 2, 3, 2)
    assert "3 batches x 2 slots from host memory" in p.stdout and "kernel launches" in p.stdout
