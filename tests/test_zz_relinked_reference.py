"""The reference's own programs on the GPU: rtlsdr_ft8d (the daemon) and ft8_lib's decode_ft8, patched as INTEGRATION.md
describes and relinked against libft8b200.so in the build container (`make -C oracle relinked` -> oracle/_ref/*_relinked, test
artefacts that travel to the GPU box like the prebuilt oracle).  Their main() is the reference's; only the hot path is ours."""
import os
import subprocess
import wave

import numpy as np
import pytest

from conftest import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DAEMON = os.path.join(ROOT, "oracle", "_ref", "rtlsdr_ft8d_relinked")
DECODE = os.path.join(ROOT, "oracle", "_ref", "decode_ft8_relinked")

needs_programs = pytest.mark.skipif(not (os.path.exists(DAEMON) and os.path.exists(DECODE)),
                                    reason="relinked reference programs not built (make -C oracle relinked needs /root/reference)")


def check_daemon_selftest(daemon, tmp_path, env=None):
    """`rtlsdr_ft8d -t` (main :1181-1190 -> decoderSelfTest :913-972 -> printSpots :643-663)."""
    # dial frequency, call sign and locator are mandatory even for the self-test (main :1157-1173); README.md:39 uses these
    p = subprocess.run([daemon, "-f", "2m", "-c", "A1XYZ", "-l", "AB12cd", "-t"], cwd=tmp_path, capture_output=True, text=True, timeout=300, env=env)
    assert p.returncode == 0, p.stderr
    assert "Self-test SUCCESS!" in p.stdout
    rows = [l.split() for l in p.stdout.splitlines() if "K1JT" in l]
    assert rows and rows[0][-2:] == ["K1JT", "FN20"], p.stdout
    assert 144174000 <= int(rows[0][1]) <= 144174100   # printSpots adds the dial frequency (2 m: 144.174 MHz) to the spot's offset
    assert os.path.getsize(os.path.join(tmp_path, "selftest.iq")) == 384000  # the reference's own writeRawIQfile ran


def check_decode_ft8(decode, tmp_path, env=None):
    """ft8_lib's decode_ft8 main() (:226-409): same stdout as the unmodified program on three of the reference's real recordings."""
    g = golden("recordings_12k")
    for k, pcm in enumerate(g["pcm"]):
        path = os.path.join(tmp_path, f"rec{k}.wav")
        with wave.open(path, "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(12000)
            w.writeframes(np.ascontiguousarray(pcm, np.int16).tobytes())
        p = subprocess.run([decode, path], capture_output=True, text=True, timeout=300, env=env)
        assert p.returncode == 0, p.stderr
        assert p.stdout.splitlines() == str(g["lines"][k]).split("\n"), str(g["names"][k])


@pytest.mark.gpu
@needs_programs
def test_reference_daemon_selftest_on_the_library(tmp_path):
    """The daemon's own main()/decoderSelfTest()/printSpots(), its initFFTW and ft8_subsystem ours, on the B200."""
    check_daemon_selftest(DAEMON, str(tmp_path))


@pytest.mark.gpu
@needs_programs
def test_reference_decode_ft8_on_the_library(tmp_path):
    """decode_ft8's own main(), its monitor_* / ft8_find_sync / ft8_decode ours, on the B200."""
    check_decode_ft8(DECODE, str(tmp_path))


@pytest.mark.gpu
@needs_programs
def test_reference_decode_ft8_on_the_library_deferred_monitor(tmp_path):
    """The same binary with FT8B200_MONITOR_DEFERRED=1: monitor_process() only appends its block, ft8_find_sync() transforms all
    93 blocks in one launch (one copy and one synchronisation per recording instead of 93).  Same stdout."""
    check_decode_ft8(DECODE, str(tmp_path), dict(os.environ, FT8B200_MONITOR_DEFERRED="1"))


@needs_programs
def test_relinked_programs_flow_with_a_cpu_stand_in(tmp_path):
    """No GPU: the SAME two binaries, with `libft8b200.so` resolved to a CPU stand-in built on the oracle
    (tests/support/cpu_stand_in.c; the binaries carry a RUNPATH, which LD_LIBRARY_PATH precedes).  This checks the patch logic and
    the expectations of the two GPU tests above (arguments, stdout, files, exit codes) in the build container; it says nothing
    about the CUDA path."""
    shim_dir = tmp_path / "standin"
    shim_dir.mkdir()
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "restate", "-s"])
    subprocess.check_call(["gcc", "-O2", "-std=gnu17", "-fPIC", "-shared", "-o", str(shim_dir / "libft8b200.so"),
                           os.path.join(ROOT, "tests", "support", "cpu_stand_in.c"), "-I", os.path.join(ROOT, "oracle"), "-I", os.path.join(ROOT, "include"),
                           "-L", os.path.join(ROOT, "oracle"), "-lft8oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-lm"])
    env = dict(os.environ, LD_LIBRARY_PATH=str(shim_dir))
    resolved = subprocess.check_output(["ldd", DAEMON], text=True, env=env)
    assert str(shim_dir) in resolved, "the stand-in must be the library the program loads in this test"
    work = tmp_path / "run"
    work.mkdir()
    check_daemon_selftest(DAEMON, str(work), env)
    check_decode_ft8(DECODE, str(work), env)
