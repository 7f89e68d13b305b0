"""The C-ABI surface without a GPU: the library loads, exports every symbol include/ft8b200.h declares,
its struct layouts are the reference's, and it fails loudly -- never silently falls
back -- when no CUDA device is usable."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ft8b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    text = re.sub(r"typedef\s+(struct|enum)\s*\{.*?\}\s*\w+\s*;", "", text, flags=re.S)
    text = re.sub(r"struct\s+\w+\s*\{.*?\}\s*;", "", text, flags=re.S)
    names = re.findall(r"\b([A-Za-z_]\w*)\s*\([^;{}]*\)\s*;", text)
    return sorted(set(names))


def test_header_declares_the_reference_entry_points():
    names = declared_functions()
    for must in ("rtlsdr_callback", "ft8_subsystem", "ft8_find_sync", "ft8_decode", "monitor_init", "monitor_process", "monitor_reset",
                 "monitor_free", "waterfall_init", "waterfall_free", "initFFTW", "freeFFTW"):
        assert must in names
    assert len(names) >= 35


def test_library_exports_every_declared_symbol(pkg):
    lib_path = pkg.LIB_PATH
    if not os.path.exists(lib_path):
        pkg.build()
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [n for n in declared_functions() if n not in exported]
    assert not missing, f"declared in include/ft8b200.h but not exported: {missing}"
    L = pkg.lib()
    assert b"sm_100a" in L.ft8b200_version()


def test_struct_layouts_match_the_reference(pkg):
    # sizes/offsets verified against the reference on x86-64
    assert C.sizeof(pkg.WaterfallT) == 40 and pkg.WaterfallT.mag.offset == 24 and pkg.WaterfallT.block_stride.offset == 32 and pkg.WaterfallT.protocol.offset == 36
    assert pkg.cand_dtype.itemsize == 8 and pkg.msg_dtype.itemsize == 28 and pkg.msg_dtype.fields["hash"][1] == 26
    assert pkg.status_dtype.itemsize == 12 and pkg.result_dtype.itemsize == 28
    assert pkg.result_dtype.fields["loc"][1] == 13 and pkg.result_dtype.fields["freq"][1] == 20 and pkg.result_dtype.fields["snr"][1] == 24
    assert C.sizeof(pkg.MonitorConfig) == 24
    assert pkg.MonitorT.wf.offset == 40 and C.sizeof(pkg.MonitorT) == 104  # same as the reference's monitor_t (decode_ft8.c:94-109)
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_mon.so")):
        ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_mon.so"))
        assert ref.refmon_sizeof_monitor() == C.sizeof(pkg.MonitorT)


def test_default_config_mirrors_the_daemon_constants(pkg):
    cfg = pkg.Config()
    pkg.lib().ft8b200_default_config(C.byref(cfg))
    # K_MAX_CANDIDATES 120, K_MAX_MESSAGES 50, K_MIN_SCORE 10, K_LDPC_ITERS 20 (rtlsdr_ft8d.h:45-48)
    assert (cfg.max_candidates, cfg.max_messages, cfg.min_score, cfg.ldpc_iterations) == (120, 50, 10, 20)


def test_fails_loudly_without_a_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.Ft8Error) as e:
        pkg.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_sources_do_not_touch_the_oracle():
    """The product path must not include, link or call anything under oracle/."""
    src = os.path.join(ROOT, "rtlsdr-ft8d_b200")
    walk = list(os.walk(src)) + list(os.walk(os.path.join(ROOT, "host"))) + list(os.walk(os.path.join(ROOT, "include")))
    walk.append((ROOT, [], ["ft8b200_loader.py"]))
    for dirpath, _, files in walk:
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".c", ".py")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                for line in text.splitlines():
                    code = line.split("//")[0]
                    assert "#include" not in code or "oracle" not in code, (f, line)
                    assert "pyoracle" not in code and "libft8oracle" not in code and "libref_" not in code, (f, line)
    so = os.path.join(src, "libft8b200.so")
    ldd = subprocess.check_output(["ldd", so], text=True)
    assert "oracle" not in ldd and "libft8oracle" not in ldd and "libref_" not in ldd
    # ... nor be the oracle-backed CPU stand-in of tests/support/ (a test artefact that must never reach the product tree):
    # the product exports its CUDA launch counter and carries sm_100a device code, the stand-in carries the oracle's symbols
    syms = subprocess.check_output(["nm", "-D", "--defined-only", so], text=True)
    assert " orc_" not in syms and "cpu_stand_in" not in syms, "an oracle-backed library sits where the product belongs"
    assert b"cpu_stand_in" not in open(so, "rb").read()
    elf_cubins = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True)
    if elf_cubins.returncode == 0:
        assert "sm_100a" in elf_cubins.stdout, "libft8b200.so carries no sm_100a device code"
    for dirpath, _, files in os.walk(src):
        for f in files:
            assert f != "cpu_stand_in.c" and not f.startswith("libft8oracle"), os.path.join(dirpath, f)


def _sass_of(kernel):
    so = os.path.join(ROOT, "rtlsdr-ft8d_b200", "libft8b200.so")
    r = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump not available")
    keep, body = False, []
    for line in r.stdout.splitlines():
        if "Function :" in line:
            keep = kernel in line
        elif keep:
            body.append(line)
    return body


def test_fir_kernel_keeps_both_roundings_of_a_tap():
    """cic_comb_fir_kernel runs a FIR tap as packed f32x2 instructions over the {I, Q} pair; the reference rounds the product AND
    the sum (rtlsdr_ft8d.c:179-192, no FMA).  ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even under -fmad=false
    (tools/f32x2_contract_probe.cu), so the kernel spells a tap as FFMA2(w, c, +0) + FADD2: the built SASS must have exactly that
    shape -- every FFMA2 with the zero register as its addend, as many FADD2 -- or the GPU's samples are no longer the reference's."""
    body = _sass_of("cic_comb_fir_kernel")
    assert body, "cic_comb_fir_kernel not found in the library"
    ffma2 = [l for l in body if " FFMA2 " in l]
    fadd2 = [l for l in body if " FADD2 " in l]
    assert len(ffma2) == 57 * 4 and len(fadd2) == 57 * 4, (len(ffma2), len(fadd2))
    assert all(l.split(";")[0].rstrip().endswith("RZ.F32") for l in ffma2), "a FIR tap was contracted into a fused multiply-add"
    assert not any(" FMUL2 " in l for l in body)


def test_fir_window_placement_is_conflict_free():
    """The {I, Q} samples' shared-memory placement in cic_comb_fir_kernel (16-byte chunk c at c ^ ((c >> 3) & 1), csrc/decimator.cu
    swz()): a permutation of the tile's chunks, and for every window position the eight 128-bit loads of a quarter-warp (thread
    stride 32 bytes) fall into eight different 16-byte bank groups; so do the comb's 64-bit stores of a half-warp."""
    swz = lambda c: c ^ ((c >> 3) & 1)
    n_chunks = (512 + 56) // 2
    assert sorted(swz(c) for c in range(n_chunks)) == list(range(n_chunks))
    for t0 in range(0, 128, 8):
        for v in range(30):
            assert len({swz(2 * (t0 + i) + v) % 8 for i in range(8)}) == 8
    for i0 in range(0, 568 - 15, 16):
        words = {(2 * swz(idx >> 1) + (idx & 1)) % 16 for idx in range(i0, i0 + 16)}
        assert len(words) == 16


def test_bad_arguments_are_rejected_before_any_launch(pkg):
    L = pkg.lib()
    assert L.ft8b200_create(None) in (None, 0) or True  # may succeed on a GPU box; must not crash
    cfg = pkg.Config(0, 1, 0, 50, 10, 20)
    assert not L.ft8b200_create(C.byref(cfg))
    assert b"bad configuration" in L.ft8b200_last_error()


_NULL_SWEEP = r'''
import ctypes as C, re, sys
hdr = open(sys.argv[1]).read()
hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
L = C.CDLL(sys.argv[2])
decls = re.findall(r"\n(int|uint64_t|uint32_t|void \*|const char \*)\s*(ft8b200_\w+)\s*\(([^;{]*?)\)\s*;", hdr)
n = 0
for ret, name, args in decls:
    params = [a.strip() for a in args.replace("\n", " ").split(",")]
    first = params[0]
    if not re.match(r"(const )?ft8b200_(ctx|pipe|cluster)_t \*", first):
        continue
    fn = getattr(L, name)
    argv = []
    for a in params:
        if "*" in a: argv.append(C.c_void_p(None))
        elif "size_t" in a or "uint64_t" in a: argv.append(C.c_uint64(0))
        elif "float" in a: argv.append(C.c_float(0.0))
        elif "double" in a: argv.append(C.c_double(0.0))
        else: argv.append(C.c_int(0))
    fn.restype = {"int": C.c_int, "uint64_t": C.c_uint64, "uint32_t": C.c_uint32, "void *": C.c_void_p, "const char *": C.c_char_p}[ret]
    print("call", name, flush=True)
    r = fn(*argv)
    if ret == "int" and not (name.endswith("_depth") or name.endswith("_in_flight") or name.endswith("_devices") or name.endswith("_nccl_version")):
        assert r < 0, (name, r)
    elif ret in ("uint64_t", "uint32_t"):
        assert r == 0, (name, r)
    elif ret == "void *":
        assert not r, (name, r)
    n += 1
print("swept", n)
'''


def test_every_entry_rejects_a_null_handle(tmp_path):
    """Every ft8b200_* entry that takes a context / pipe / cluster handle, called with that handle NULL and every other argument
    zero: an error code (or 0 / NULL for the counters and getters), never a crash and never a CUDA call -- runs without a GPU.
    The signatures are read from include/ft8b200.h, so a new entry point is swept as soon as it is declared."""
    script = tmp_path / "sweep.py"
    script.write_text(_NULL_SWEEP)
    so = os.path.join(ROOT, "rtlsdr-ft8d_b200", "libft8b200.so")
    p = subprocess.run([sys.executable, str(script), os.path.join(ROOT, "include", "ft8b200.h"), so], capture_output=True, text=True, timeout=120)
    last = [l for l in p.stdout.splitlines() if l.startswith("call")][-1:] or ["(none)"]
    assert p.returncode == 0, f"crashed or wrong answer at {last[0]}: rc {p.returncode}\n{p.stderr[-600:]}"
    swept = int(p.stdout.strip().splitlines()[-1].split()[1])
    assert swept >= 70, swept


_ZERO_SWEEP = _NULL_SWEEP.replace("print(\"swept\", n)", "").replace('''    r = fn(*argv)
    if ret == "int" and not (name.endswith("_depth") or name.endswith("_in_flight") or name.endswith("_devices") or name.endswith("_nccl_version")):
        assert r < 0, (name, r)
    elif ret in ("uint64_t", "uint32_t"):
        assert r == 0, (name, r)
    elif ret == "void *":
        assert not r, (name, r)
    n += 1''', '''    kind = re.match(r"(const )?ft8b200_(ctx|pipe|cluster)_t", first).group(2)
    if kind == "cluster" or name in ("ft8b200_device_malloc",):
        continue
    argv[0] = C.c_void_p(handles[kind])
    r = fn(*argv)
    n += 1''').replace("n = 0\n", '''n = 0
L.ft8b200_create.restype = C.c_void_p
L.ft8b200_pipe_create.restype = C.c_void_p
L.ft8b200_last_error.restype = C.c_char_p
handles = {"ctx": L.ft8b200_create(None), "pipe": L.ft8b200_pipe_create(None, 2)}
assert handles["ctx"] and handles["pipe"], L.ft8b200_last_error()
''') + '''
# the context still works: one silent slot through the whole path
import numpy as np
zi = np.zeros(48000, np.float32); res = np.zeros(50 * 28, np.uint8); nres = C.c_int32(-7)
rc = L.ft8b200_process_slots_host(C.c_void_p(handles["ctx"]), zi.ctypes.data_as(C.c_void_p), zi.ctypes.data_as(C.c_void_p), 1, res.ctypes.data_as(C.c_void_p), C.byref(nres))
assert rc == 0 and nres.value == 0, (rc, nres.value, L.ft8b200_last_error())
print("swept", n)
'''


@pytest.mark.gpu
def test_every_entry_survives_zero_arguments_on_a_live_handle(tmp_path):
    """The same sweep on a B200 with LIVE context and pipe handles and every other argument zero / NULL: whatever each entry
    answers, the process does not crash, nothing is launched on garbage, and the context decodes a slot afterwards."""
    script = tmp_path / "sweep0.py"
    script.write_text(_ZERO_SWEEP)
    so = os.path.join(ROOT, "rtlsdr-ft8d_b200", "libft8b200.so")
    p = subprocess.run([sys.executable, str(script), os.path.join(ROOT, "include", "ft8b200.h"), so], capture_output=True, text=True, timeout=300)
    last = [l for l in p.stdout.splitlines() if l.startswith("call")][-1:] or ["(none)"]
    assert p.returncode == 0, f"crashed at {last[0]}: rc {p.returncode}\n{p.stdout[-300:]}\n{p.stderr[-900:]}"
    assert int(p.stdout.strip().splitlines()[-1].split()[1]) >= 55


def test_file_readers_survive_fuzzed_files(tmp_path):
    """The on-disk readers on 3000 hostile files per seed (random bytes, RIFF headers with every field out of range, truncated and
    oversized data): no crash, no heap damage (glibc aborts at exit on a smashed heap).  The fuzz found one: a header claiming mono
    16-bit with blockAlign 4 made fread() write 4 bytes per sample into the caller's 2-byte-per-sample buffer."""
    for seed in (1, 2):
        p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_file_readers.py"), str(seed)], capture_output=True, text=True, timeout=300)
        assert p.returncode == 0 and "fuzzed 3000" in p.stdout, (p.returncode, p.stderr[-500:])


def test_wav_with_wrong_block_align_reads_like_the_reference(pkg, tmp_path):
    """blockAlign 4 in a mono 16-bit header: load_wav() reads n * 4 bytes and takes the first n int16 of them (wave.c:104-121); the
    library does the same and writes exactly n samples into the caller's buffers."""
    import struct
    n, align = 500, 4
    data = np.arange(n * align // 2, dtype=np.int16) - 300
    hdr = b"RIFF" + struct.pack("<I", 36 + n * align) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, 12000, 24000, align, 16) + b"data" + struct.pack("<I", n * align)
    path = tmp_path / "align4.wav"
    path.write_bytes(hdr + data.tobytes())
    L = pkg.lib()
    raw = np.full(n + 64, 0x5A5A, np.int16)
    sig = np.full(n + 64, 7.0, np.float32)
    ns, sr = C.c_int(n + 64), C.c_int(0)
    assert L.ft8b200_load_wav_s16(raw.ctypes.data_as(C.c_void_p), sig.ctypes.data_as(C.c_void_p), C.byref(ns), C.byref(sr), str(path).encode()) == 0
    assert ns.value == n and sr.value == 12000
    assert np.array_equal(raw[:n], data[:n]) and (raw[n:] == 0x5A5A).all(), "exactly n samples, the first n int16 of the data"
    assert np.array_equal(sig[:n], data[:n].astype(np.float32) / np.float32(32768.0)) and (sig[n:] == 7.0).all()


def test_host_entries_survive_fuzzed_arguments():
    """pack77 on arbitrary byte strings and the three report builders on records without terminators, at every output capacity
    from 0 up (tools/fuzz_host_entries.py; clean under ASan/UBSan too, profiles/sanitizer_r2.md): no crash, no heap damage."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_host_entries.py"), "1"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "fuzzed 4000" in p.stdout, (p.returncode, p.stderr[-500:])
