/* oracle/ref_mon_harness.c -- ft8_lib's example decoder (decode_ft8.c) compiled
 * unmodified as a shared object: the 12 kHz real-audio waterfall path
 * (monitor_init/process/reset/free, waterfall_init/free: decode_ft8.c:63-224, kiss_fftr).
 * TEST INFRASTRUCTURE ONLY.  monitor_t / monitor_config_t are defined inside the .c
 * (decode_ft8.c:82-109), so the accessors below are the only portable way to reach them. */
#define main ref_decode_ft8_main
#include "decode_ft8.c"
#undef main

/* linked with -Wl,--wrap=malloc: every malloc() in the reference objects returns zeroed memory, so the uninitialised
 * last_frame of monitor_init() (decode_ft8.c:131) is deterministic, also inside the reference's own main() */
void *__wrap_malloc(size_t n) { return calloc(1, n); }

int refmon_sizeof_monitor(void) { return (int)sizeof(monitor_t); }
monitor_t *refmon_new(float f_min, float f_max, int sample_rate, int time_osr, int freq_osr, int protocol) {
    monitor_config_t cfg = { f_min, f_max, sample_rate, time_osr, freq_osr, (ftx_protocol_t)protocol };
    monitor_t *m = (monitor_t *)calloc(1, sizeof(monitor_t));
    monitor_init(m, &cfg);
    /* the reference leaves last_frame uninitialised (decode_ft8.c:131); zero it so the oracle is deterministic */
    memset(m->last_frame, 0, sizeof(float) * (size_t)m->nfft);
    return m;
}
void refmon_delete(monitor_t *m) { monitor_free(m); free(m); }
void refmon_process(monitor_t *m, const float *frame) { monitor_process(m, frame); }
void refmon_reset(monitor_t *m) { monitor_reset(m); }
int refmon_info(monitor_t *m, int *out /* block_size, subblock_size, nfft, max_blocks, num_blocks, num_bins, time_osr, freq_osr, block_stride */) {
    out[0] = m->block_size; out[1] = m->subblock_size; out[2] = m->nfft; out[3] = m->wf.max_blocks;
    out[4] = m->wf.num_blocks; out[5] = m->wf.num_bins; out[6] = m->wf.time_osr; out[7] = m->wf.freq_osr;
    out[8] = m->wf.block_stride;
    return 9;
}
const uint8_t *refmon_mag(monitor_t *m) { return m->wf.mag; }
const float *refmon_window(monitor_t *m) { return m->window; }
float refmon_max_mag(monitor_t *m) { return m->max_mag; }
int refmon_find_sync(monitor_t *m, int num_candidates, candidate_t *heap, int min_score) {
    return ft8_find_sync(&m->wf, num_candidates, heap, min_score);
}
int refmon_decode(monitor_t *m, const candidate_t *cand, int max_iters, message_t *msg, decode_status_t *status) {
    memset(status, 0xA5, sizeof(*status));
    memset(msg, 0, sizeof(*msg));
    return ft8_decode(&m->wf, cand, msg, max_iters, status) ? 1 : 0;
}
/* raw kiss_fftr of one real frame: golden vectors for the 3840-point real FFT */
void refmon_fftr(int nfft, const float *timedata, float *freq_ri /* 2*(nfft/2+1) */) {
    kiss_fftr_cfg cfg = kiss_fftr_alloc(nfft, 0, NULL, NULL);
    kiss_fftr(cfg, timedata, (kiss_fft_cpx *)freq_ri);
    free(cfg);
}
int refmon_load_wav(const char *path, float *signal, int *num_samples, int *sample_rate) {
    return load_wav(signal, num_samples, sample_rate, path);
}
