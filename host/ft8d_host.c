/* ft8d_host.c -- the C host side of the path: what rtlsdr_ft8d's main()/decoder thread and ft8_lib's decode_ft8 do
 * around the hot path, written against libft8b200.so's C ABI only (include/ft8b200.h; no CUDA headers, no C++).
 *
 * It is the reference-language caller of the drop-in entry points, so the boundary is exercised from plain C the way a
 * maintainer's patched daemon would exercise it:
 *
 *   selftest [OUT.iq]        the daemon's `-t` flow (rtlsdr_ft8d.c:913-972,1181-1190): the "CQ K1JT FN20" known-answer
 *                            message -> tones -> FSK at 3200 sps + noise -> initFFTW / ft8_subsystem -> spot table
 *   decode  F.iq|F.c2 ...    the daemon's `-r` flow (decodeRecordedFile, :859-887), any number of files in ONE batch
 *   receive RAW.u8           the live flow (:76-285,1336-1354): raw 2.4 Msps uint8 IQ handed to rtlsdr_callback() in
 *                            librtlsdr-sized buffers, the 15 s buffer flip, decoder() on every closed slot
 *   wav     [-ft4] F.wav ... ft8_lib's decode_ft8 main() (decode_ft8.c:226-409) through monitor_* / ft8_find_sync /
 *                            ft8_decode, printing the same lines
 *   batch                    B200-side extra: synthetic raw slots through the pipelined executor from host memory
 *   cluster [-n devices] [streams [slots_per_stream]]   B200-side extra: BASELINE config #5 in miniature over every visible GPU from this
 *                            one process: receiver streams made on their own device, sharded by stream, decoded-spot records
 *                            gathered with NCCL (ft8b200_cluster_t); prints one line per (stream, slot) -- the same lines whatever
 *                            the number of devices
 *   latency SLOT.iq [RAW.u8 [F.wav]]   single-slot latency of the literal drop-in calls, one JSON line (BASELINE config #1; the
 *                            reference publishes this as "decode burst per 15 s slot", README.md:153-157)
 *
 * Exit status: 0 = ran (selftest: and decoded the known answer), 1 = self-test failed / bad usage, 2 = library error.
 * Without a usable B200 the library reports why and the program stops: there is no CPU path behind these calls.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "ft8b200.h"

#define SLOT FT8B200_SLOT_SAMPLES
#define MAX_MESSAGES 50      /* K_MAX_MESSAGES, rtlsdr_ft8d.h:47 */
#define MAX_CANDIDATES 120   /* kMax_candidates, decode_ft8.c:21 */
#define MIN_SCORE 10         /* kMin_score, decode_ft8.c:20 */
#define LDPC_ITERATIONS 20   /* kLDPC_iterations, decode_ft8.c:22 */

static uint32_t g_dial_hz = 14074000u; /* -f: dec_options.freq */
static uint32_t g_unixtime = 0;        /* -T: fixed time stamp for reproducible "No spot" lines */

static int fail(const char *what) {
    fprintf(stderr, "ft8d_host: %s: %s\n", what, ft8b200_last_error());
    return 2;
}

static uint32_t stamp(uint32_t back_seconds) {
    return g_unixtime ? g_unixtime : (uint32_t)time(NULL) - back_seconds + 1u;
}

/* printSpots() (rtlsdr_ft8d.c:635-663) through the library's formatter */
static void print_spots(const struct decoder_results *spots, int32_t n, uint32_t unixtime) {
    char table[4096];
    if (ft8b200_format_spots(spots, n < 0 ? 0u : (uint32_t)n, g_dial_hz, unixtime, table, sizeof table) >= 0) fputs(table, stdout);
}

/* ---------------------------------------------------------------------------------------------- selftest */

/* 64-bit LCG + Box-Muller: the reference draws from rand(), which differs between C libraries; this one is the same everywhere */
static uint64_t g_rng = 0x9E3779B97F4A7C15ull;
static double uniform01(void) {
    g_rng = g_rng * 6364136223846793005ull + 1442695040888963407ull;
    return (double)((g_rng >> 11) + 1u) / 9007199254740993.0;
}
static float gaussian(float sigma) {
    const double r = sqrt(-2.0 * log(uniform01())), a = 6.283185307179586 * uniform01();
    return (float)(r * cos(a)) * sigma;
}

/* writeRawIQfile() (rtlsdr_ft8d.c:784-808): 48000 pairs of float32, interleaved I, -Q */
static int write_iq_file(const char *path, const float *rail_i, const float *rail_q) {
    FILE *f = fopen(path, "wb");
    if (!f) { perror(path); return -1; }
    for (int k = 0; k < SLOT; ++k) {
        const float pair[2] = {rail_i[k], -rail_q[k]};
        if (fwrite(pair, sizeof pair, 1, f) != 1) { fclose(f); return -1; }
    }
    return fclose(f);
}

static int run_selftest(const char *save_as) {
    static const uint8_t kat_payload[10] = {0x00, 0x00, 0x00, 0x20, 0x4d, 0xfc, 0xdc, 0x8a, 0x14, 0x08};
    static const char kat_tones[] = "3140652000000001005477547106035036373140652547441342116056460065174427143140652";
    static float rail_i[SLOT], rail_q[SLOT];
    static struct decoder_results spots[MAX_MESSAGES];
    uint8_t payload[10], tones[105];

    if (ft8b200_pack77("CQ K1JT FN20QI", payload) != 0) return fail("pack77"); /* the reference's own test message, :918 */
    if (memcmp(payload, kat_payload, 10) != 0) {
        fprintf(stderr, "ft8d_host: packed message differs from the known answer (rtlsdr_ft8d.c:919-923)\n");
        return 1;
    }
    ft8b200_config_t cfg;
    ft8b200_default_config(&cfg);
    cfg.max_slots = 1;
    ft8b200_ctx_t *ctx = ft8b200_create(&cfg);
    if (!ctx) return fail("ft8b200_create");
    if (ft8b200_encode_tones(ctx, payload, 1, PROTO_FT8, tones) != 0) return fail("encode_tones");
    ft8b200_destroy(ctx);
    for (int k = 0; k < 79; ++k)
        if (tones[k] != (uint8_t)(kat_tones[k] - '0')) {
            fprintf(stderr, "ft8d_host: channel symbol %d differs from the known answer\n", k);
            return 1;
        }

    /* 79 symbols of 512 samples from t = 0, tone 0 at 50 Hz - 3.5 bins, amplitude 0.5 over sigma 0.02 noise per rail */
    const double bin_hz = 3200.0 / 512.0, two_pi_dt = 6.283185307179586 / 3200.0;
    double phase = 0.0;
    for (int sym = 0; sym < 79; ++sym) {
        const double step = two_pi_dt * (50.0 + ((double)tones[sym] - 3.5) * bin_hz);
        for (int j = 0; j < 512; ++j, phase += step) {
            rail_i[512 * sym + j] = 0.5f * (float)cos(phase) + gaussian(0.02f);
            rail_q[512 * sym + j] = 0.5f * (float)sin(phase) + gaussian(0.02f);
        }
    }

    if (save_as && write_iq_file(save_as, rail_i, rail_q) != 0) return 1; /* the reference always leaves "selftest.iq" behind */

    int32_t n = 0;
    initFFTW();
    ft8_subsystem(rail_i, rail_q, SLOT, spots, &n);
    freeFFTW();
    print_spots(spots, n, stamp(0));
    if (n >= 1 && strcmp(spots[0].call, "K1JT") == 0 && strcmp(spots[0].loc, "FN20") == 0) {
        puts("Self-test SUCCESS!");
        return 0;
    }
    fputs("Self-test FAILED!\n", stderr);
    return 1;
}

/* ---------------------------------------------------------------------------------------------- decode (.iq / .c2) */

static int run_decode(int n_files, char **paths) {
    if (n_files < 1) return 1;
    ft8b200_config_t cfg;
    ft8b200_default_config(&cfg);
    cfg.max_slots = n_files;
    ft8b200_ctx_t *ctx = ft8b200_create(&cfg);
    if (!ctx) return fail("ft8b200_create");
    struct decoder_results *spots = calloc((size_t)n_files * MAX_MESSAGES, sizeof *spots);
    int32_t *counts = calloc((size_t)n_files, sizeof *counts), *samples = calloc((size_t)n_files, sizeof *samples);
    if (!spots || !counts || !samples) return 2;
    if (ft8b200_decode_iq_files(ctx, (const char *const *)paths, n_files, spots, counts, samples) != 0) return fail("decode_iq_files");
    for (int k = 0; k < n_files; ++k) {
        const size_t len = strlen(paths[k]);
        if (len < 3 || (strcmp(paths[k] + len - 3, ".iq") != 0 && strcmp(paths[k] + len - 3, ".c2") != 0)) {
            fprintf(stderr, "Not a valid extension!! (only .iq & .c2 files)\n");
            continue;
        }
        printf("Number of samples: %d\n", samples[k]);
        if (samples[k]) print_spots(spots + (size_t)k * MAX_MESSAGES, counts[k], stamp(120));
    }
    free(spots); free(counts); free(samples);
    ft8b200_destroy(ctx);
    return 0;
}

/* ---------------------------------------------------------------------------------------------- receive (raw IQ) */

/* Plays a recording of the dongle's byte stream the way librtlsdr's async reader delivers it: `chunk` bytes per
 * rtlsdr_callback() with ctx = NULL (the reference's registration, rtlsdr_ft8d.c:214), the 15 s flip between two
 * callbacks once a slot's worth of bytes has gone by (main()'s timer, :1339-1354), then decoder() (:221-285). */
static int run_receive(const char *path, uint32_t chunk, int own_stream) {
    FILE *f = strcmp(path, "-") == 0 ? stdin : fopen(path, "rb");
    if (!f) { perror(path); return 1; }
    unsigned char *buf = malloc(chunk);
    static struct decoder_results spots[MAX_MESSAGES];
    if (!buf) return 2;
    ft8b200_ctx_t *ctx = NULL;
    ft8b200_stream_t *rx = NULL; /* NULL = the process-wide stream */
    if (own_stream) {
        ft8b200_config_t cfg;
        ft8b200_default_config(&cfg);
        cfg.max_slots = 1;
        if (!(ctx = ft8b200_create(&cfg)) || !(rx = ft8b200_stream_create(ctx))) return fail("stream_create");
    } else {
        initFFTW();
    }
    uint64_t fed = 0, next_flip = FT8B200_RAW_SLOT_BYTES;
    int slot = 0, rc = 0, eof = 0;
    while (!eof && !rc) {
        size_t got = fread(buf, 1, chunk, f);
        eof = got < chunk;
        got -= got % 8u; /* the mixer consumes 4 complex samples at a time (:129) */
        if (got) {
            rtlsdr_callback(buf, (uint32_t)got, rx);
            fed += got;
        }
        /* the timer fires between two callbacks; at the end of the recording the slot being filled is closed too */
        while (!rc && (fed >= next_flip || (eof && fed + FT8B200_RAW_SLOT_BYTES > next_flip))) {
            int32_t n = 0;
            const uint32_t have = ft8b200_stream_count(rx);
            if (ft8b200_stream_flip(rx) != 0 || ft8b200_stream_decode(rx, spots, &n) != 0) { rc = fail("stream_decode"); break; }
            if (n < 0) printf("slot %d: %u samples, signal too short, skipped\n", slot, have);
            else { printf("slot %d: %u samples\n", slot, have); print_spots(spots, n, stamp(15)); }
            ++slot;
            next_flip += FT8B200_RAW_SLOT_BYTES;
        }
    }
    if (own_stream) { ft8b200_stream_destroy(rx); ft8b200_destroy(ctx); }
    else freeFFTW();
    free(buf);
    if (f != stdin) fclose(f);
    return rc;
}

/* ---------------------------------------------------------------------------------------------- wav (decode_ft8) */

static int run_wav_file(const char *path, ftx_protocol_t protocol) {
    static float audio[15 * 12000];
    int sample_rate = 12000, n_samples = 15 * 12000;
    if (ft8b200_load_wav(audio, &n_samples, &sample_rate, path) < 0) return -1;

    monitor_config_t mc = {100.0f, 3000.0f, sample_rate, 2, 2, protocol};
    monitor_t mon;
    monitor_init(&mon, &mc);
    for (int at = 0; at + mon.block_size <= n_samples; at += mon.block_size) monitor_process(&mon, audio + at);

    candidate_t cand[MAX_CANDIDATES];
    const int n_cand = ft8_find_sync(&mon.wf, MAX_CANDIDATES, cand, MIN_SCORE);

    /* first-seen unique messages: open addressing on the CRC, keyed by (hash, text) like decode_ft8.c:336-393 */
    message_t seen[MAX_MESSAGES];
    bool used[MAX_MESSAGES] = {false};
    int n_unique = 0;
    for (int c = 0; c < n_cand && n_unique < MAX_MESSAGES; ++c) {
        if (cand[c].score < MIN_SCORE) continue;
        message_t msg;
        decode_status_t status;
        if (!ft8_decode(&mon.wf, &cand[c], &msg, LDPC_ITERATIONS, &status)) continue;
        int at = msg.hash % MAX_MESSAGES;
        while (used[at] && !(seen[at].hash == msg.hash && strcmp(seen[at].text, msg.text) == 0)) at = (at + 1) % MAX_MESSAGES;
        if (used[at]) continue; /* duplicate */
        seen[at] = msg;
        used[at] = true;
        ++n_unique;
        const float freq_hz = (cand[c].freq_offset + (float)cand[c].freq_sub / mon.wf.freq_osr) / mon.symbol_period;
        const float time_sec = (cand[c].time_offset + (float)cand[c].time_sub / mon.wf.time_osr) * mon.symbol_period;
        printf("000000 %3d %+4.2f %4.0f ~  %s\n", cand[c].score, time_sec, freq_hz, msg.text);
    }
    monitor_free(&mon);
    return n_unique;
}

static int run_wav(int argc, char **argv) {
    ftx_protocol_t protocol = PROTO_FT8;
    int files = 0;
    for (int k = 0; k < argc; ++k) {
        if (strcmp(argv[k], "-ft4") == 0) { protocol = PROTO_FT4; continue; }
        if (files++) printf("-- %s\n", argv[k]);
        const int n = run_wav_file(argv[k], protocol);
        if (n < 0) { fprintf(stderr, "ft8d_host: cannot read %s\n", argv[k]); return 1; }
        fprintf(stderr, "Decoded %d messages\n", n);
    }
    return files ? 0 : 1;
}

/* ---------------------------------------------------------------------------------------------- batch (executor) */

/* n_batches batches of n_slots synthetic raw slots from HOST memory through ft8b200_pipe_t: the end-to-end call a
 * multi-receiver host makes. */
static int run_batch(int n_slots, int n_batches, int depth) {
    ft8b200_config_t cfg;
    ft8b200_default_config(&cfg);
    cfg.max_slots = n_slots;
    ft8b200_pipe_t *pipe = ft8b200_pipe_create(&cfg, depth);
    if (!pipe) return fail("pipe_create");
    /* The host only owns bytes here: noise-only recordings, which decode to no spot.  What this exercises from C is the
     * executor's bookkeeping (submit until EBUSY-free, collect in order, counts); parity of contents is the tests' job. */
    const size_t bytes = (size_t)FT8B200_RAW_SLOT_BYTES;
    uint8_t *iq = malloc(bytes * (size_t)n_slots);
    struct decoder_results *spots = calloc((size_t)n_slots * MAX_MESSAGES, sizeof *spots);
    int32_t *counts = calloc((size_t)n_slots, sizeof *counts);
    if (!iq || !spots || !counts) return 2;
    for (size_t k = 0; k < bytes * (size_t)n_slots; ++k) {
        g_rng = g_rng * 6364136223846793005ull + 1442695040888963407ull;
        iq[k] = (uint8_t)(128 + (int)((g_rng >> 40) & 63u) - 32);
    }
    int submitted = 0, collected = 0, rc = 0;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    while (collected < n_batches) {
        while (submitted < n_batches && ft8b200_pipe_in_flight(pipe) < ft8b200_pipe_depth(pipe)) {
            if ((rc = ft8b200_pipe_submit_host(pipe, iq, bytes, n_slots)) != 0) break;
            ++submitted;
        }
        if (rc) break;
        const int got = ft8b200_pipe_collect(pipe, spots, counts, n_slots);
        if (got != n_slots) { rc = got < 0 ? got : FT8B200_EINVAL; break; }
        ++collected;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (rc) { fprintf(stderr, "ft8d_host: executor: %s\n", ft8b200_pipe_error(pipe)); return 2; }
    const double sec = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    printf("%d batches x %d slots from host memory: %.1f slots/s, %llu kernel launches\n", n_batches, n_slots,
           (double)n_batches * n_slots / sec, (unsigned long long)ft8b200_pipe_kernel_launches(pipe));
    ft8b200_pipe_destroy(pipe);
    free(iq); free(spots); free(counts);
    return 0;
}

/* ---------------------------------------------------------------------------------------------- cluster (config #5) */

/* Stream s, slot g carries "CQ <call(s, g)> <grid>" at its own frequency; everything is derived from the GLOBAL stream index, so the
 * bytes -- and therefore the spots -- do not depend on which device a stream lands on. */
static void cluster_call(int stream, int slot, char *call7, char *grid5) {
    const unsigned k = (unsigned)(stream * 8 + slot);
    call7[0] = 'K'; call7[1] = (char)('0' + k % 10u); call7[2] = (char)('A' + (k / 10u) % 26u); call7[3] = (char)('A' + (k / 260u) % 26u);
    call7[4] = (char)('A' + (k * 7u) % 26u); call7[5] = 0;
    grid5[0] = (char)('A' + (k * 5u) % 18u); grid5[1] = (char)('A' + (k * 11u) % 18u); grid5[2] = (char)('0' + k % 10u); grid5[3] = (char)('0' + (k / 3u) % 10u);
    grid5[4] = 0;
}

static int run_cluster(int n_devices, int n_streams, int slots_per_stream) {
    ft8b200_config_t cfg;
    ft8b200_default_config(&cfg);
    ft8b200_cluster_t *cl = ft8b200_cluster_create(&cfg, n_devices, 2);
    if (!cl) return fail("cluster_create");
    const int nd = ft8b200_cluster_devices(cl);
    const size_t slot_bytes = FT8B200_RAW_SLOT_BYTES, stream_bytes = slot_bytes * (size_t)slots_per_stream;
    uint8_t *d_iq[64] = {0};
    int counts[64] = {0};
    if (nd > 64) return 1;
    for (int d = 0; d < nd; ++d) {
        int first = 0, count = 0;
        ft8b200_cluster_shard(cl, n_streams, d, &first, &count);
        counts[d] = count;
        if (!count) continue;
        ft8b200_ctx_t *ctx = ft8b200_cluster_ctx(cl, d);
        const int rows = count * slots_per_stream;
        d_iq[d] = ft8b200_device_malloc(ctx, stream_bytes * (size_t)count);
        ft8b200_signal_t *sig = calloc((size_t)rows, sizeof *sig);
        int *firsts = calloc((size_t)rows + 1, sizeof *firsts);
        if (!d_iq[d] || !sig || !firsts) { fprintf(stderr, "ft8d_host: out of memory\n"); return 2; }
        for (int r = 0; r < rows; ++r) {
            const int stream = first + r / slots_per_stream, slot = r % slots_per_stream;
            char call[8], grid[8];
            cluster_call(stream, slot, call, grid);
            if (ft8b200_pack77_std("CQ", call, grid, sig[r].payload) != 0) { fprintf(stderr, "ft8d_host: cannot pack CQ %s %s\n", call, grid); return 1; }
            sig[r].f0_hz = 300.0f + 37.0f * (float)((stream * 8 + slot) % 29);
            sig[r].t0_sec = 0.5f;
            sig[r].amp = 20.0f;
            firsts[r + 1] = r + 1;
        }
        /* rows of one stream are consecutive in memory: the stream is one continuous recording of slots_per_stream slots */
        if (ft8b200_synth_raw(ctx, sig, firsts, rows, 30.0f, 0xC5u, first * slots_per_stream, d_iq[d], slot_bytes, slot_bytes, NULL) != 0) return fail("synth_raw");
        free(sig); free(firsts);
    }
    const int total = n_streams * slots_per_stream;
    struct decoder_results *spots = calloc((size_t)total * MAX_MESSAGES, sizeof *spots);
    int32_t *n_spots = calloc((size_t)total, sizeof *n_spots);
    if (!spots || !n_spots) return 2;
    if (ft8b200_cluster_submit_streams(cl, (const uint8_t *const *)d_iq, stream_bytes, stream_bytes, counts, slots_per_stream, slot_bytes) != 0 ||
        ft8b200_cluster_collect(cl, spots, n_spots, total) != total) {
        fprintf(stderr, "ft8d_host: cluster: %s\n", ft8b200_cluster_error(cl));
        return 2;
    }
    int good = 0;
    for (int r = 0; r < total; ++r) {
        char call[8], grid[8];
        cluster_call(r / slots_per_stream, r % slots_per_stream, call, grid);
        const struct decoder_results *s = spots + (size_t)r * MAX_MESSAGES;
        printf("stream %d slot %d: %d message(s)", r / slots_per_stream, r % slots_per_stream, n_spots[r]);
        for (int k = 0; k < n_spots[r] && k < MAX_MESSAGES; ++k)
            if (s[k].call[0]) printf("  %s %s %d Hz", s[k].call, s[k].loc, s[k].freq);
        putchar('\n');
        for (int k = 0; k < n_spots[r] && k < MAX_MESSAGES; ++k)
            if (strcmp(s[k].call, call) == 0 && strcmp(s[k].loc, grid) == 0) { ++good; break; }
    }
    fprintf(stderr, "%d devices, %d of %d slots decoded to their own message, %llu NCCL gather(s) (NCCL %d), %llu kernel launches\n", nd, good, total,
            (unsigned long long)ft8b200_cluster_gathers(cl), ft8b200_cluster_nccl_version(cl), (unsigned long long)ft8b200_cluster_kernel_launches(cl));
    for (int d = 0; d < nd; ++d) if (d_iq[d]) ft8b200_device_free(ft8b200_cluster_ctx(cl, d), d_iq[d]);
    ft8b200_cluster_destroy(cl);
    free(spots); free(n_spots);
    return good == total ? 0 : 1;
}

/* ---------------------------------------------------------------------------------------------- latency (config #1) */

static double now_ms(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return 1e3 * (double)t.tv_sec + 1e-6 * (double)t.tv_nsec;
}
static int cmp_double(const void *a, const void *b) { return (*(const double *)a > *(const double *)b) - (*(const double *)a < *(const double *)b); }
static double median(double *v, int n) { qsort(v, (size_t)n, sizeof *v, cmp_double); return v[n / 2]; }

/* One 15 s slot at a time through the reference's own entry points, wall-clock per call as the daemon's decoder thread sees it:
 *   subsystem_ms   decoder()'s conditioning on the host (rtlsdr_ft8d.c:242-263) + ft8_subsystem(I, Q, 48000, spots, &n)
 *   receive_ms     1099 x rtlsdr_callback(65536 bytes) of one raw 2.4 Msps slot + the 15 s flip + decoder() (stream_decode)
 *   wav_ms         decode_ft8's main() flow: 93 x monitor_process + ft8_find_sync(120) + one ft8_decode per candidate
 *   wav_deferred_ms  the same flow with ft8b200_monitor_set_deferred(&mon, 1): one waterfall launch instead of 93 round trips */
static int run_latency(const char *iq_path, const char *raw_path, const char *wav_path, int reps) {
    static float rail_i[SLOT], rail_q[SLOT], ci[SLOT], cq[SLOT];
    static struct decoder_results spots[MAX_MESSAGES];
    double t[64];
    if (reps < 3) reps = 3;
    if (reps > 64) reps = 64;
    float peak = 0.0f;
    const int n_pairs = ft8b200_read_iq_file(iq_path, rail_i, rail_q, &peak);
    if (n_pairs <= 0) { fprintf(stderr, "ft8d_host: cannot read %s\n", iq_path); return 1; }
    initFFTW();
    int32_t n = 0, n_sub = 0;
    for (int r = 0; r < reps + 2; ++r) {   /* two untimed calls first (context creation, first launches) */
        const double t0 = now_ms();
        float max_sig = 1e-24f;            /* decoder(): zero tail is already there; peak-normalise to 0.5 */
        for (int k = 0; k < SLOT; ++k) {
            const float a = fabsf(rail_i[k]), b = fabsf(rail_q[k]);
            if (a > max_sig) max_sig = a;
            if (b > max_sig) max_sig = b;
        }
        const float scale = (float)(0.5 / (double)max_sig);
        for (int k = 0; k < SLOT; ++k) { ci[k] = rail_i[k] * scale; cq[k] = rail_q[k] * scale; }
        ft8_subsystem(ci, cq, SLOT, spots, &n);
        if (r >= 2) t[r - 2] = now_ms() - t0;
        n_sub = n;
    }
    const double subsystem_ms = median(t, reps);

    double receive_ms = -1.0;
    int32_t n_rx = -1;
    if (raw_path && raw_path[0] && strcmp(raw_path, "-") != 0) {
        FILE *f = fopen(raw_path, "rb");
        unsigned char *raw = malloc(FT8B200_RAW_SLOT_BYTES);
        if (!f || !raw || fread(raw, 1, FT8B200_RAW_SLOT_BYTES, f) != FT8B200_RAW_SLOT_BYTES) { fprintf(stderr, "ft8d_host: cannot read one raw slot from %s\n", raw_path); return 1; }
        fclose(f);
        const int rx_reps = reps < 5 ? reps : 5;
        for (int r = 0; r < rx_reps + 1; ++r) {
            const double t0 = now_ms();
            for (size_t at = 0; at < FT8B200_RAW_SLOT_BYTES; at += 65536u) {
                size_t len = FT8B200_RAW_SLOT_BYTES - at;
                if (len > 65536u) len = 65536u;
                rtlsdr_callback(raw + at, (uint32_t)(len - len % 8u), NULL);
            }
            if (ft8b200_stream_flip(NULL) != 0 || ft8b200_stream_decode(NULL, spots, &n) != 0) return fail("stream_decode");
            if (r >= 1) t[r - 1] = now_ms() - t0;
            n_rx = n;
        }
        receive_ms = median(t, rx_reps);
        free(raw);
    }

    double wav_ms = -1.0, wav_deferred_ms = -1.0;
    int n_wav = -1, n_wav_deferred = -1;
    for (int deferred = 0; deferred < 2 && wav_path && wav_path[0]; ++deferred) {
        static float audio[15 * 12000];
        int sample_rate = 12000, n_samples = 15 * 12000;
        if (ft8b200_load_wav(audio, &n_samples, &sample_rate, wav_path) < 0) { fprintf(stderr, "ft8d_host: cannot read %s\n", wav_path); return 1; }
        /* the monitor is set up once, like a daemon would: what is timed is one recording's worth of calls */
        monitor_config_t mc = {100.0f, 3000.0f, sample_rate, 2, 2, PROTO_FT8};
        monitor_t mon;
        monitor_init(&mon, &mc);
        if (deferred) ft8b200_monitor_set_deferred(&mon, 1);   /* monitor_process() appends; ft8_find_sync() transforms all 93 blocks in one launch */
        for (int r = 0; r < reps + 1; ++r) {
            const double t0 = now_ms();
            monitor_reset(&mon);
            for (int at = 0; at + mon.block_size <= n_samples; at += mon.block_size) monitor_process(&mon, audio + at);
            candidate_t cand[MAX_CANDIDATES];
            const int n_cand = ft8_find_sync(&mon.wf, MAX_CANDIDATES, cand, MIN_SCORE);
            int ok = 0;
            for (int c = 0; c < n_cand; ++c) {
                message_t msg;
                decode_status_t status;
                ok += ft8_decode(&mon.wf, &cand[c], &msg, LDPC_ITERATIONS, &status) ? 1 : 0;
            }
            if (r >= 1) t[r - 1] = now_ms() - t0;
            if (deferred) n_wav_deferred = ok; else n_wav = ok;
        }
        monitor_free(&mon);
        if (deferred) wav_deferred_ms = median(t, reps); else wav_ms = median(t, reps);
    }
    freeFFTW();
    printf("{\"subsystem_ms\": %.4f, \"subsystem_results\": %d, \"receive_ms\": %.4f, \"receive_results\": %d, \"wav_ms\": %.4f, \"wav_decodes\": %d, "
           "\"wav_deferred_ms\": %.4f, \"wav_deferred_decodes\": %d, \"reps\": %d}\n",
           subsystem_ms, n_sub, receive_ms, n_rx, wav_ms, n_wav, wav_deferred_ms, n_wav_deferred, reps);
    return 0;
}

/* ---------------------------------------------------------------------------------------------- main */

static int usage(void) {
    fputs("usage: ft8d_host [-f dial_hz] [-T unixtime] selftest [save_as.iq]\n"
          "       ft8d_host [-f dial_hz] [-T unixtime] decode file.iq|file.c2 ...\n"
          "       ft8d_host [-f dial_hz] [-T unixtime] receive [-b bytes_per_callback] [-s] raw_iq.u8|-\n"
          "       ft8d_host wav [-ft4] file.wav ...\n"
          "       ft8d_host batch [slots [batches [depth]]]\n"
          "       ft8d_host cluster [-n devices] [streams [slots_per_stream]]\n"
          "       ft8d_host latency slot.iq [raw_slot.u8|- [file.wav [reps]]]\n",
          stderr);
    return 1;
}

int main(int argc, char **argv) {
    int a = 1;
    for (; a + 1 < argc && argv[a][0] == '-' && argv[a][1] && !argv[a][2]; a += 2) {
        if (argv[a][1] == 'f') g_dial_hz = (uint32_t)strtoul(argv[a + 1], NULL, 10);
        else if (argv[a][1] == 'T') g_unixtime = (uint32_t)strtoul(argv[a + 1], NULL, 10);
        else return usage();
    }
    if (a >= argc) return usage();
    const char *cmd = argv[a++];
    if (strcmp(cmd, "selftest") == 0) return run_selftest(argc > a ? argv[a] : NULL);
    if (strcmp(cmd, "decode") == 0) return argc > a ? run_decode(argc - a, argv + a) : usage();
    if (strcmp(cmd, "wav") == 0) return run_wav(argc - a, argv + a);
    if (strcmp(cmd, "receive") == 0) {
        uint32_t chunk = 65536u;
        int own = 0;
        for (; a < argc && argv[a][0] == '-' && argv[a][1]; ++a) {
            if (strcmp(argv[a], "-s") == 0) own = 1;
            else if (strcmp(argv[a], "-b") == 0 && a + 1 < argc) chunk = (uint32_t)strtoul(argv[++a], NULL, 10);
            else return usage();
        }
        if (a >= argc || chunk < 8u || chunk % 8u) return usage();
        return run_receive(argv[a], chunk, own);
    }
    if (strcmp(cmd, "cluster") == 0) {
        int nd = 0;
        if (a + 1 < argc && strcmp(argv[a], "-n") == 0) { nd = atoi(argv[a + 1]); a += 2; }
        const int streams = argc > a ? atoi(argv[a]) : 4, spp = argc > a + 1 ? atoi(argv[a + 1]) : 2;
        return streams > 0 && spp > 0 && spp <= 8 ? run_cluster(nd, streams, spp) : usage();
    }
    if (strcmp(cmd, "latency") == 0)
        return argc > a ? run_latency(argv[a], argc > a + 1 ? argv[a + 1] : NULL, argc > a + 2 ? argv[a + 2] : NULL, argc > a + 3 ? atoi(argv[a + 3]) : 15) : usage();
    if (strcmp(cmd, "batch") == 0) {
        const int slots = argc > a ? atoi(argv[a]) : 4, batches = argc > a + 1 ? atoi(argv[a + 1]) : 4, depth = argc > a + 2 ? atoi(argv[a + 2]) : 2;
        return slots > 0 && batches > 0 && depth > 0 ? run_batch(slots, batches, depth) : usage();
    }
    return usage();
}
