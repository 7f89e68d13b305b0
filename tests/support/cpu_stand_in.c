/* cpu_stand_in.c -- TEST INFRASTRUCTURE.  A CPU stand-in for the drop-in layer of libft8b200.so, built on the oracle
 * (oracle/libft8oracle.so), so that the reference's own programs patched per INTEGRATION.md can be run END TO END in a
 * container without a GPU: it checks the patch logic and the expectations of tests/test_zz_relinked_reference.py (arguments,
 * stdout, files, exit codes).  It is linked only by tests/test_integration_link.py, under the library's soname in a temp
 * directory; the product never sees it, and it proves nothing about the CUDA path -- the GPU tests do that. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ft8_oracle.h"
#include "ft8b200.h"

/* marker the no-oracle guard (tests/test_abi.py::test_product_sources_do_not_touch_the_oracle) rejects by name */
const char ft8b200_cpu_stand_in_marker[] = "cpu_stand_in: oracle-backed test artefact, never the product";

void initFFTW(void) {}
void freeFFTW(void) {}

void ft8_subsystem(float *iSamples, float *qSamples, uint32_t samples_len, struct decoder_results *decodes, int32_t *n_results) {
    (void)samples_len;
    static orc_slot_report_t rep;
    orc_result_t res[50];
    memset(res, 0, sizeof res);
    *n_results = orc_subsystem(iSamples, qSamples, 120, 50, 10, 20, res, &rep, NULL, NULL);
    memcpy(decodes, res, sizeof res); /* same 28-byte layout */
}

/* ---- the process-wide receiver stream (ctx == NULL) ---- */
static orc_decim_t g_dec;
static int g_dec_ready = 0, g_buf = 0;
static float g_i[2][48000], g_q[2][48000];
static size_t g_n[2];

void rtlsdr_callback(unsigned char *samples, uint32_t samples_count, void *ctx) {
    (void)ctx;
    if (!g_dec_ready) { orc_decim_reset(&g_dec); g_dec_ready = 1; }
    orc_decim_feed(&g_dec, samples, samples_count, g_i[g_buf], g_q[g_buf], NULL, NULL, 48000, &g_n[g_buf]);
}
int ft8b200_stream_flip(ft8b200_stream_t *s) { (void)s; g_buf ^= 1; g_n[g_buf] = 0; return 0; }
uint32_t ft8b200_stream_count(ft8b200_stream_t *s) { (void)s; return (uint32_t)g_n[g_buf]; }
int ft8b200_stream_fetch(ft8b200_stream_t *s, float *h_i, float *h_q, uint32_t *n_valid) {
    (void)s;
    const int prev = g_buf ^ 1;
    memset(h_i, 0, sizeof(float) * 48000); memset(h_q, 0, sizeof(float) * 48000);
    memcpy(h_i, g_i[prev], sizeof(float) * g_n[prev]); memcpy(h_q, g_q[prev], sizeof(float) * g_n[prev]);
    if (n_valid) *n_valid = (uint32_t)g_n[prev];
    return 0;
}
int ft8b200_stream_decode(ft8b200_stream_t *s, struct decoder_results *h_results, int32_t *h_nresults) {
    static float ci[48000], cq[48000];
    uint32_t n = 0;
    ft8b200_stream_fetch(s, ci, cq, &n);
    if (n < (15 - 3) * 3200) { *h_nresults = -1; return 0; }
    orc_condition(ci, cq, n, 48000);
    ft8_subsystem(ci, cq, 48000, h_results, h_nresults);
    return 0;
}

/* ---- ft8_lib's monitor + sync + decode ---- */
void monitor_init(monitor_t *me, const monitor_config_t *cfg) {
    memset(me, 0, sizeof *me);
    orc_monitor_t *m = orc_monitor_new(cfg->sample_rate, cfg->time_osr, cfg->freq_osr, (int)cfg->protocol);
    int info[9];
    orc_monitor_info(m, info); /* block_size, subblock_size, nfft, max_blocks, num_blocks, num_bins, time_osr, freq_osr, block_stride */
    me->symbol_period = cfg->protocol == PROTO_FT4 ? 0.048f : 0.160f;
    me->block_size = info[0]; me->subblock_size = info[1]; me->nfft = info[2];
    me->wf.max_blocks = info[3]; me->wf.num_blocks = info[4]; me->wf.num_bins = info[5];
    me->wf.time_osr = info[6]; me->wf.freq_osr = info[7]; me->wf.block_stride = info[8];
    me->wf.protocol = cfg->protocol;
    me->wf.mag = (uint8_t *)orc_monitor_mag(m);
    me->fft_work = m;
}
void monitor_process(monitor_t *me, const float *frame) {
    orc_monitor_t *m = (orc_monitor_t *)me->fft_work;
    orc_monitor_process(m, frame);
    int info[9];
    orc_monitor_info(m, info);
    me->wf.num_blocks = info[4];
    me->max_mag = orc_monitor_max_mag(m);
}
void monitor_reset(monitor_t *me) { orc_monitor_reset((orc_monitor_t *)me->fft_work); me->wf.num_blocks = 0; me->max_mag = 0; }
void monitor_free(monitor_t *me) { orc_monitor_free((orc_monitor_t *)me->fft_work); me->fft_work = NULL; }

int ft8_find_sync(const waterfall_t *power, int num_candidates, candidate_t heap[], int min_score) {
    return orc_find_sync((const orc_waterfall_t *)power, num_candidates, (orc_candidate_t *)heap, min_score);
}
bool ft8_decode(const waterfall_t *power, const candidate_t *cand, message_t *message, int max_iterations, decode_status_t *status) {
    return orc_decode((const orc_waterfall_t *)power, (const orc_candidate_t *)cand, max_iterations, (orc_message_t *)message,
                      (orc_status_t *)status, NULL, NULL) != 0;
}
