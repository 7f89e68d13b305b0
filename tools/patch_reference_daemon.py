#!/usr/bin/env python3
"""The reference-side binding of INTEGRATION.md section 1, applied mechanically: a PRIVATE, patched copy of the reference daemon
(rtlsdr_ft8d.c / rtlsdr_ft8d.h) in which the hot path comes from libft8b200.so instead of the daemon's own code.

    python tools/patch_reference_daemon.py /root/reference OUT_DIR

writes OUT_DIR/rtlsdr_ft8d.c and OUT_DIR/rtlsdr_ft8d.h (ft8_lib is used where it lies).  Nothing is written into the repository
or into the reference tree, and no reference source is kept: tests/test_integration_link.py builds the copy in a temp directory
to prove that the binding compiles and links (and that the resulting daemon stops loudly on a machine without a B200).

The patch, in the order INTEGRATION.md lists it:
  1. rtlsdr_ft8d.h:144   `static void rtlsdr_callback(...)`  ->  `void rtlsdr_callback(...)`   (the definition lives in the library)
  2. rtlsdr_ft8d.c       the definitions of rtlsdr_callback (:76-202), initFFTW (:314-336), freeFFTW (:338-347) and
                         ft8_subsystem (:1387-1524) are removed
  3. rtlsdr_ft8d.c       after the ft8_lib includes:  #define FT8B200_WITH_RTLSDR_FT8D_H / #include "ft8b200.h"
  4. main() (:1349-1351) the 15 s buffer flip also closes the slot on the device: ft8b200_stream_flip(NULL)
  5. decoder() (:235-278) "too short" test, tail clearing, normalisation and ft8_subsystem() call -> ft8b200_stream_decode(NULL, ...)
     (+ ft8b200_stream_fetch(NULL, ...) when the daemon was asked to write the samples to a file, so that saveSample() :280 still works)
and for ft8_lib's example decoder: OUT_DIR/decode_ft8_main.c = ft8_lib/decode_ft8.c with lines 35-224
(window functions, waterfall_*, monitor_*) replaced by `#include "ft8b200.h"`.
"""
import os
import re
import sys


def remove_function(src: str, signature_regex: str) -> str:
    """Cut `<signature> { ... }` (brace-matched, comments and strings of this file contain no unbalanced braces) out of src."""
    m = re.search(signature_regex, src, flags=re.M)
    if not m:
        raise SystemExit("patch_reference_daemon: signature not found: " + signature_regex)
    i = src.index("{", m.end() - 1)
    depth, k = 0, i
    while True:
        c = src[k]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                break
        k += 1
    return src[:m.start()] + "/* (provided by libft8b200.so) */\n" + src[k + 1:]


def patch(ref_dir: str, out_dir: str):
    c = open(os.path.join(ref_dir, "rtlsdr_ft8d.c")).read()
    h = open(os.path.join(ref_dir, "rtlsdr_ft8d.h")).read()

    # 1
    assert "static void rtlsdr_callback(" in h
    h = h.replace("static void rtlsdr_callback(", "void rtlsdr_callback(")

    # 2
    c = remove_function(c, r"^static void rtlsdr_callback\(unsigned char \*samples, uint32_t samples_count, void \*ctx\) \{")
    c = remove_function(c, r"^void initFFTW\(\) \{")
    c = remove_function(c, r"^void freeFFTW\(\) \{")
    c = remove_function(c, r"^void ft8_subsystem\(float \*iSamples,[^{]*\{")

    # 3
    anchor = '#include "./ft8_lib/ft8/encode.h"\n'
    assert anchor in c
    c = c.replace(anchor, anchor + '\n#define FT8B200_WITH_RTLSDR_FT8D_H\n#include "ft8b200.h"\n', 1)

    # 4
    flip = "        rx_state.iqIndex[rx_state.bufferIndex] = 0;\n"
    assert c.count(flip) == 1
    c = c.replace(flip, flip + "        ft8b200_stream_flip(NULL);  /* close the slot on the device, start the next one */\n")

    # 5: from the "too short" test up to and including the ft8_subsystem() call of decoder()
    start = c.index("        if (rx_state.iqIndex[prevBuffer] < ( (SIGNAL_LENGHT - 3) * SIGNAL_SAMPLE_RATE ) ) {")
    call = c.index("        ft8_subsystem(rx_state.iSamples[prevBuffer],", start)
    end = c.index(");", call) + 2
    block = c[start:end]
    assert "unixtime = unixtime - 15 + 1;" in block
    replacement = (
        "        (void)prevBuffer;\n"
        "        /* decoder() of the library: skip a slot shorter than 12 s, clear the tail, normalise, search & decode */\n"
        "        if (ft8b200_stream_decode(NULL, dec_results, &n_results) != 0 || n_results < 0) {\n"
        "            LOG(LOG_DEBUG, \"Decoder thread -- Signal too short, skipping!\\n\");\n"
        "            n_results = 0;\n"
        "            continue;\n"
        "        }\n"
        "        rx_options.nloop++;\n"
        "        time_t unixtime;\n"
        "        time ( &unixtime );\n"
        "        unixtime = unixtime - 15 + 1;\n"
        "        rx_state.gtm = gmtime( &unixtime );\n"
        "        if (rx_options.writefile)  /* saveSample() below writes the slot's samples: bring them to the host */\n"
        "            ft8b200_stream_fetch(NULL, rx_state.iSamples[prevBuffer], rx_state.qSamples[prevBuffer], &rx_state.iqIndex[prevBuffer]);\n")
    c = c[:start] + replacement + c[end:]

    os.makedirs(out_dir, exist_ok=True)
    open(os.path.join(out_dir, "rtlsdr_ft8d.c"), "w").write(c)
    open(os.path.join(out_dir, "rtlsdr_ft8d.h"), "w").write(h)


def patch_decode_ft8(ref_dir: str, out_dir: str):
    """ft8_lib/decode_ft8.c with its own window functions, waterfall_* and monitor_* (lines 35-224)
    replaced by `#include "ft8b200.h"`; main() (lines 226-409) is the reference's, untouched."""
    c = open(os.path.join(ref_dir, "ft8_lib", "decode_ft8.c")).read()
    start = c.index("static float hann_i(int i, int N)")
    end = c.index("int main(int argc, char** argv)")
    assert "void monitor_reset(monitor_t* me)" in c[start:end] and "void waterfall_init(" in c[start:end]
    c = c[:start] + '#include "ft8b200.h" /* waterfall_init/free, monitor_init/process/reset/free, monitor_t, monitor_config_t */\n\n' + c[end:]
    os.makedirs(out_dir, exist_ok=True)
    open(os.path.join(out_dir, "decode_ft8_main.c"), "w").write(c)


if __name__ == "__main__":
    if len(sys.argv) != 3:
        raise SystemExit(__doc__)
    patch(sys.argv[1], sys.argv[2])
    patch_decode_ft8(sys.argv[1], sys.argv[2])
