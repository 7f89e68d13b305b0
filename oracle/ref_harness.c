/* oracle/ref_harness.c -- builds the UNMODIFIED reference daemon into a shared
 * object with stage taps.  TEST INFRASTRUCTURE ONLY: nothing under oracle/ is
 * imported, linked or executed by the product path (rtlsdr-ft8d_b200/csrc).
 *
 * How it works (no reference source is copied into this repository):
 *   - the reference translation unit is textually included from where it lies
 *     under /root/reference (REF root given with -I), with `main` renamed so the
 *     file-static rtlsdr_callback (rtlsdr_ft8d.c:76) becomes reachable;
 *   - while it is being included, ft8_find_sync / ft8_decode are renamed to
 *     tap_* so that the calls made by ft8_subsystem() (rtlsdr_ft8d.c:1450,1476)
 *     land in the recorders below, which forward to the real ft8_lib functions;
 *   - ft8_lib/ft8/decode.c is compiled with -Dbp_decode=tap_bp_decode so the
 *     normalised LLRs and LDPC hard decisions of every candidate are recorded.
 * The FFT is the reference's own vendored kiss_fft behind oracle/shims/fftw3.h.
 */
#define main ref_daemon_main
#define ft8_find_sync tap_find_sync
#define ft8_decode tap_decode
#include "rtlsdr_ft8d.c"
#undef ft8_decode
#undef ft8_find_sync
#undef main

int ft8_find_sync(const waterfall_t *power, int num_candidates, candidate_t heap[], int min_score);
bool ft8_decode(const waterfall_t *power, const candidate_t *cand, message_t *message, int max_iterations, decode_status_t *status);

#define TAP_MAX_CAND 1024
#define TAP_WF_BYTES (93 * 2 * 2 * 960) /* big enough for the 12 kHz monitor waterfall too */

typedef struct {
    int32_t wf_bytes;                 /* bytes valid in wf[] */
    int32_t wf_dims[6];               /* num_blocks,num_bins,time_osr,freq_osr,block_stride,protocol */
    int32_t n_cand;                   /* returned by ft8_find_sync */
    int32_t n_decode_calls;           /* ft8_decode calls seen since reset */
    int32_t n_bp_calls;
    candidate_t cand[TAP_MAX_CAND];   /* sorted list as returned */
    candidate_t dec_cand[TAP_MAX_CAND];
    int32_t dec_ok[TAP_MAX_CAND];
    decode_status_t dec_status[TAP_MAX_CAND]; /* pre-filled with 0xA5 bytes: unwritten fields stay visible */
    message_t dec_msg[TAP_MAX_CAND];
    float llr[TAP_MAX_CAND][FTX_LDPC_N];      /* input of bp_decode (after ftx_normalize_logl) */
    uint8_t plain[TAP_MAX_CAND][FTX_LDPC_N];
    int32_t bp_errors[TAP_MAX_CAND];
    uint8_t wf[TAP_WF_BYTES];
} ref_taps_t;

static ref_taps_t g_taps;

/* Scripted mode: ft8_subsystem()'s own candidate loop / duplicate table / CQ filter (rtlsdr_ft8d.c:1452-1523) run over a
 * HAND-MADE candidate list and hand-made decode results, so the a15 edge cases (hash clashes with different text, 2-token
 * CQ messages, many unique messages) are answered by the reference's code itself.  While armed, the two taps below hand
 * out the script instead of calling ft8_lib. */
static struct {
    int armed, n;
    const candidate_t *heap_base;
    candidate_t cand[TAP_MAX_CAND];
    int32_t ok[TAP_MAX_CAND];
    message_t msg[TAP_MAX_CAND];
} g_script;

void *ref_taps_ptr(void) { return &g_taps; }
int ref_taps_size(void) { return (int)sizeof(g_taps); }
void ref_taps_reset(void) { memset(&g_taps, 0, sizeof(g_taps)); }

int tap_find_sync(const waterfall_t *power, int num_candidates, candidate_t heap[], int min_score) {
    int nbytes = power->num_blocks * power->block_stride;
    if (nbytes > TAP_WF_BYTES) nbytes = TAP_WF_BYTES;
    memcpy(g_taps.wf, power->mag, (size_t)nbytes);
    g_taps.wf_bytes = nbytes;
    g_taps.wf_dims[0] = power->num_blocks; g_taps.wf_dims[1] = power->num_bins;
    g_taps.wf_dims[2] = power->time_osr;   g_taps.wf_dims[3] = power->freq_osr;
    g_taps.wf_dims[4] = power->block_stride; g_taps.wf_dims[5] = (int)power->protocol;
    if (g_script.armed) {
        int n = g_script.n < num_candidates ? g_script.n : num_candidates;
        memcpy(heap, g_script.cand, sizeof(candidate_t) * (size_t)n);
        g_script.heap_base = heap;
        g_taps.n_cand = n;
        return n;
    }
    int n = ft8_find_sync(power, num_candidates, heap, min_score);
    g_taps.n_cand = n;
    for (int i = 0; i < n && i < TAP_MAX_CAND; ++i) g_taps.cand[i] = heap[i];
    return n;
}

bool tap_decode(const waterfall_t *power, const candidate_t *cand, message_t *message, int max_iterations, decode_status_t *status) {
    int i = g_taps.n_decode_calls;
    memset(status, 0xA5, sizeof(*status));
    memset(message, 0, sizeof(*message));
    if (g_script.armed) {
        const long k = cand - g_script.heap_base;
        g_taps.n_decode_calls = i + 1;
        memset(status, 0, sizeof(*status));
        if (k < 0 || k >= g_script.n || !g_script.ok[k]) { status->ldpc_errors = 1; return false; }
        *message = g_script.msg[k];
        return true;
    }
    bool ok = ft8_decode(power, cand, message, max_iterations, status);
    if (i < TAP_MAX_CAND) {
        g_taps.dec_cand[i] = *cand;
        g_taps.dec_ok[i] = ok ? 1 : 0;
        g_taps.dec_status[i] = *status;
        g_taps.dec_msg[i] = *message;
    }
    g_taps.n_decode_calls = i + 1;
    return ok;
}

void bp_decode(float codeword[], int max_iters, uint8_t plain[], int *ok);
void tap_bp_decode(float codeword[], int max_iters, uint8_t plain[], int *ok) {
    int i = g_taps.n_bp_calls;
    if (i < TAP_MAX_CAND) memcpy(g_taps.llr[i], codeword, sizeof(float) * FTX_LDPC_N);
    bp_decode(codeword, max_iters, plain, ok);
    if (i < TAP_MAX_CAND) {
        memcpy(g_taps.plain[i], plain, FTX_LDPC_N);
        g_taps.bp_errors[i] = *ok;
    }
    g_taps.n_bp_calls = i + 1;
}

/* ---- daemon-level entry points ------------------------------------------------ */
int ref_k_max_candidates(void) { return K_MAX_CANDIDATES; }
int ref_k_max_messages(void) { return K_MAX_MESSAGES; }
int ref_sizeof_results(void) { return (int)sizeof(struct decoder_results); }

static int g_inited = 0;
void ref_init(void) {
    if (!g_inited) { initFFTW(); g_inited = 1; }
    initSampleStorage();
    memset(dec_results, 0, sizeof(dec_results));
}

/* rtlsdr_callback (rtlsdr_ft8d.c:76-202). The reference mutates `buf` in place. */
void ref_callback(unsigned char *buf, uint32_t nbytes) { rtlsdr_callback(buf, nbytes, NULL); }
uint32_t ref_rx_count(int which) { return rx_state.iqIndex[which]; }
uint32_t ref_rx_buffer_index(void) { return rx_state.bufferIndex; }
void ref_rx_copy(int which, float *i_out, float *q_out, uint32_t n) {
    memcpy(i_out, rx_state.iSamples[which], n * sizeof(float));
    memcpy(q_out, rx_state.qSamples[which], n * sizeof(float));
}
/* what main() does every 15 s (rtlsdr_ft8d.c:1339-1354), minus the thread signalling */
void ref_rx_flip(void) {
    rx_state.bufferIndex = (rx_state.bufferIndex + 1) % 2;
    rx_state.iqIndex[rx_state.bufferIndex] = 0;
}

/* ft8_subsystem (rtlsdr_ft8d.c:1387-1524); dec_results is zeroed first so stale gaps are visible */
int32_t ref_subsystem(float *i_samples, float *q_samples, struct decoder_results *out, int32_t out_cap) {
    int32_t n = 0;
    ref_taps_reset();
    memset(dec_results, 0, sizeof(dec_results));
    ft8_subsystem(i_samples, q_samples, SIGNAL_LENGHT * SIGNAL_SAMPLE_RATE, dec_results, &n);
    int cap = (int)(sizeof(dec_results) / sizeof(dec_results[0]));
    if (out_cap < cap) cap = out_cap;
    memcpy(out, dec_results, sizeof(struct decoder_results) * (size_t)cap);
    return n;
}

/* ft8_subsystem()'s candidate loop over a scripted candidate list (see g_script); the waterfall it computes from the
 * silent input is ignored by the scripted taps */
int32_t ref_subsystem_scripted(const candidate_t *cand, const int32_t *ok, const message_t *msg, int32_t n, struct decoder_results *out,
                               int32_t out_cap) {
    static float zeros_i[SIGNAL_LENGHT * SIGNAL_SAMPLE_RATE], zeros_q[SIGNAL_LENGHT * SIGNAL_SAMPLE_RATE];
    if (n > TAP_MAX_CAND) n = TAP_MAX_CAND;
    memset(&g_script, 0, sizeof(g_script));
    memcpy(g_script.cand, cand, sizeof(candidate_t) * (size_t)n);
    memcpy(g_script.ok, ok, sizeof(int32_t) * (size_t)n);
    memcpy(g_script.msg, msg, sizeof(message_t) * (size_t)n);
    g_script.n = n;
    g_script.armed = 1;
    const int32_t r = ref_subsystem(zeros_i, zeros_q, out, out_cap);
    g_script.armed = 0;
    return r;
}

/* ---- wall-clock of the reference's own single-slot flows (BASELINE config #1; what README.md:153-157 calls the "decode burst") ---- */
#include <time.h>
static double ref_now_ms(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return 1e3 * (double)t.tv_sec + 1e-6 * (double)t.tv_nsec;
}
/* decoder()'s conditioning (rtlsdr_ft8d.c:242-263; the function itself is the body of a thread loop and cannot be called) */
static void ref_condition(float *i_s, float *q_s, uint32_t n_valid) {
    for (uint32_t k = n_valid; k < SIGNAL_LENGHT * SIGNAL_SAMPLE_RATE; ++k) { i_s[k] = 0.0f; q_s[k] = 0.0f; }
    float maxSig = 1e-24f;
    for (uint32_t k = 0; k < SIGNAL_LENGHT * SIGNAL_SAMPLE_RATE; ++k) {
        const float a = fabsf(i_s[k]), b = fabsf(q_s[k]);
        if (a > maxSig) maxSig = a;
        if (b > maxSig) maxSig = b;
    }
    maxSig = 0.5 / maxSig;
    for (uint32_t k = 0; k < SIGNAL_LENGHT * SIGNAL_SAMPLE_RATE; ++k) { i_s[k] *= maxSig; q_s[k] *= maxSig; }
}
/* conditioning + ft8_subsystem on one slot, `reps` times: out_ms[r]; returns n_results of the last call */
int32_t ref_time_subsystem(const float *i_samples, const float *q_samples, int reps, double *out_ms) {
    static float ci[SIGNAL_LENGHT * SIGNAL_SAMPLE_RATE], cq[SIGNAL_LENGHT * SIGNAL_SAMPLE_RATE];
    int32_t n = 0;
    for (int r = 0; r < reps; ++r) {
        const double t0 = ref_now_ms();
        memcpy(ci, i_samples, sizeof(ci));
        memcpy(cq, q_samples, sizeof(cq));
        ref_condition(ci, cq, SIGNAL_LENGHT * SIGNAL_SAMPLE_RATE);
        memset(dec_results, 0, sizeof(dec_results));
        ft8_subsystem(ci, cq, SIGNAL_LENGHT * SIGNAL_SAMPLE_RATE, dec_results, &n);
        out_ms[r] = ref_now_ms() - t0;
    }
    return n;
}
/* one raw 2.4 Msps slot: rtlsdr_callback() in 65536-byte calls, the 15 s flip, conditioning, ft8_subsystem; returns n_results.
 * `raw` is modified in place by the reference's mixer (rtlsdr_ft8d.c:129-140). */
int32_t ref_time_receive(unsigned char *raw, uint32_t nbytes, double *out_ms) {
    const double t0 = ref_now_ms();
    for (uint32_t at = 0; at < nbytes; at += 65536u) {
        uint32_t len = nbytes - at;
        if (len > 65536u) len = 65536u;
        rtlsdr_callback(raw + at, len - len % 8u, NULL);
    }
    const uint32_t which = rx_state.bufferIndex, have = rx_state.iqIndex[which];
    ref_rx_flip();
    int32_t n = 0;
    ref_condition(rx_state.iSamples[which], rx_state.qSamples[which], have);
    memset(dec_results, 0, sizeof(dec_results));
    ft8_subsystem(rx_state.iSamples[which], rx_state.qSamples[which], SIGNAL_LENGHT * SIGNAL_SAMPLE_RATE, dec_results, &n);
    *out_ms = ref_now_ms() - t0;
    return n;
}

int32_t ref_selftest(void) { ref_taps_reset(); return decoderSelfTest(); }

/* ---- ft8_lib level entry points ------------------------------------------------ */
static waterfall_t make_wf(uint8_t *mag, int num_blocks, int num_bins, int time_osr, int freq_osr, int protocol) {
    waterfall_t wf;
    wf.max_blocks = num_blocks; wf.num_blocks = num_blocks; wf.num_bins = num_bins;
    wf.time_osr = time_osr; wf.freq_osr = freq_osr; wf.mag = mag;
    wf.block_stride = time_osr * freq_osr * num_bins; wf.protocol = (ftx_protocol_t)protocol;
    return wf;
}
int ref_find_sync(uint8_t *mag, int num_blocks, int num_bins, int time_osr, int freq_osr, int protocol,
                  int num_candidates, candidate_t *heap, int min_score) {
    waterfall_t wf = make_wf(mag, num_blocks, num_bins, time_osr, freq_osr, protocol);
    return ft8_find_sync(&wf, num_candidates, heap, min_score);
}
/* returns ok; llr_out/plain_out receive the bp_decode tap of this call */
int ref_decode(uint8_t *mag, int num_blocks, int num_bins, int time_osr, int freq_osr, int protocol,
               const candidate_t *cand, int max_iters, message_t *msg, decode_status_t *status,
               float *llr_out, uint8_t *plain_out) {
    waterfall_t wf = make_wf(mag, num_blocks, num_bins, time_osr, freq_osr, protocol);
    g_taps.n_bp_calls = 0;
    memset(status, 0xA5, sizeof(*status));
    memset(msg, 0, sizeof(*msg));
    bool ok = ft8_decode(&wf, cand, msg, max_iters, status);
    if (llr_out) memcpy(llr_out, g_taps.llr[0], sizeof(float) * FTX_LDPC_N);
    if (plain_out) memcpy(plain_out, g_taps.plain[0], FTX_LDPC_N);
    return ok ? 1 : 0;
}
void ref_bp_decode(const float *llr, int max_iters, uint8_t *plain, int *errors) {
    float cw[FTX_LDPC_N];
    memcpy(cw, llr, sizeof(cw));
    bp_decode(cw, max_iters, plain, errors);
}
int ref_pack77(const char *msg, uint8_t *c77) { return pack77(msg, c77); }
void ref_encode(const uint8_t *payload, uint8_t *tones) { ft8_encode(payload, tones); }
int ref_unpack77(const uint8_t *a77, char *text) { return unpack77(a77, text); }
uint32_t ref_crc(const uint8_t *msg, int num_bits) { return ftx_compute_crc(msg, num_bits); }
void ref_add_crc(const uint8_t *payload, uint8_t *a91) { ftx_add_crc(payload, a91); }

/* read-only views of the protocol tables, used by tools/gen_tables.py to derive and
 * cross-check this repository's own tables (never at run time on the GPU box) */
const uint8_t *ref_table_nm(void) { return &kFTX_LDPC_Nm[0][0]; }
const uint8_t *ref_table_mn(void) { return &kFTX_LDPC_Mn[0][0]; }
const uint8_t *ref_table_num_rows(void) { return &kFTX_LDPC_Num_rows[0]; }
const uint8_t *ref_table_generator(void) { return &kFTX_LDPC_generator[0][0]; }
const uint8_t *ref_table_costas(void) { return &kFT8_Costas_pattern[0]; }
const uint8_t *ref_table_gray(void) { return &kFT8_Gray_map[0]; }
/* the window table built by initFFTW (rtlsdr_ft8d.c:331-334) */
const float *ref_window(void) { return hann; }
