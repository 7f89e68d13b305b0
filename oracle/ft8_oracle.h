/* oracle/ft8_oracle.h -- CPU restatement of the rtlsdr-ft8d receive-and-decode hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load libft8oracle.so.  The product
 * (rtlsdr-ft8d_b200/csrc, libft8b200.so) never includes, links or calls anything here.
 *
 * Every function restates one piece of the reference (file:line cited at each definition in
 * ft8_oracle.c) in plain scalar C with the reference's exact operation order, so that its
 * results are bit-identical to the reference compiled with
 *     gcc -O3 -std=gnu17 -ffp-contract=off -fwrapv
 * Pinning: tests/test_oracle_vs_ref.py compares every stage below against the unmodified
 * reference (oracle/_ref/libref_*.so, built from /root/reference by oracle/Makefile) on seeded
 * inputs, and tests/test_oracle_golden.py against the committed fixtures in tests/golden/
 * (generated from the reference by tools/make_golden.py) plus the reference's known-answer
 * constants (rtlsdr_ft8d.c:919-923, ft8_lib/test.c:97-101).
 * One boundary stays unpinned against the real third-party dependency: the daemon calls FFTW3f
 * (external, plan-dependent rounding); the oracle -- like oracle/_ref -- uses the reference's
 * vendored kiss_fft arithmetic instead.
 */
#ifndef FT8_ORACLE_H
#define FT8_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_SLOT_SAMPLES 48000      /* 15 s @ 3200 sps           (rtlsdr_ft8d.h:34-35)  */
#define ORC_DECIM 751               /* samples per output, NOT 750 (rtlsdr_ft8d.c:156-160) */
#define ORC_FIR_TAPS 57
#define ORC_WF_BYTES 94208          /* 92*2*2*256                (rtlsdr_ft8d.h:54)     */

/* ---- ABI mirrors (ft8_lib/ft8/decode.h:15-53, rtlsdr_ft8d.h:136-141) ---- */
typedef struct { int max_blocks, num_blocks, num_bins, time_osr, freq_osr; uint8_t *mag; int block_stride; int protocol; } orc_waterfall_t;
typedef struct { int16_t score, time_offset, freq_offset; uint8_t time_sub, freq_sub; } orc_candidate_t;
typedef struct { char text[25]; uint16_t hash; } orc_message_t;
typedef struct { int ldpc_errors; uint16_t crc_extracted, crc_calculated; int unpack_status; } orc_status_t;
typedef struct { char call[13]; char loc[7]; int32_t freq; int32_t snr; } orc_result_t;

/* ---- a1-a3: fs/4 mixer + CIC(N=2, R=751, M=2) + 57-tap FIR, streaming state ---- */
typedef struct {
    int32_t ix1, ix2, qx1, qx2;           /* integrators            */
    int32_t it1y, it1z, qt1y, qt1z;       /* comb 1 delay line      */
    int32_t it2y, it2z, qt2y, qt2z;       /* comb 2 delay line      */
    uint32_t decim_index;
    float fir_i[56], fir_q[56];           /* FIR history of (float)y2 */
    uint64_t n_out;                        /* outputs produced so far  */
} orc_decim_t;
void orc_decim_reset(orc_decim_t *st);
/* One rtlsdr_callback() worth of samples.  nbytes must be a multiple of 8.  Appends up to
 * cap-*count outputs to i_out/q_out starting at *count (the reference drops outputs once its
 * 48000-sample buffer is full but keeps filtering); y2 taps (int32, before the FIR) are written
 * to y2i/y2q when non-NULL with the same indexing. Input is NOT modified. */
void orc_decim_feed(orc_decim_t *st, const uint8_t *iq, size_t nbytes,
                    float *i_out, float *q_out, int32_t *y2i, int32_t *y2q, size_t cap, size_t *count);
const float *orc_fir_coefs(void); /* 57 floats */

/* ---- a4: decoder() conditioning ---- */
float orc_condition(float *i_s, float *q_s, size_t n_valid, size_t n_total);

/* ---- a5: daemon waterfall (1024-pt c2c, sine window, u8 dB) ---- */
void orc_sine_window(float *w, int n);
void orc_fft_c2c(int n, const float *in_ri, float *out_ri);           /* kiss_fft restated */
void orc_fft_r2c(int n, const float *in, float *out_ri);              /* kiss_fftr restated, n even */
uint8_t orc_quantize_db(float x);                                     /* x = 1e-12f + scaled |X|^2 */
void orc_db_thresholds(float *t257);                                  /* t[k] = min x with quantize(x) >= k; t[0]=0, t[256]=+inf */
void orc_waterfall_daemon(const float *i_s, const float *q_s, uint8_t *mag /*94208*/);

/* ---- a5': ft8_lib monitor (12 kHz real audio) ---- */
typedef struct orc_monitor orc_monitor_t;
orc_monitor_t *orc_monitor_new(int sample_rate, int time_osr, int freq_osr, int protocol);
void orc_monitor_free(orc_monitor_t *m);
void orc_monitor_process(orc_monitor_t *m, const float *frame);
void orc_monitor_reset(orc_monitor_t *m);
void orc_monitor_info(const orc_monitor_t *m, int *out9);
const uint8_t *orc_monitor_mag(const orc_monitor_t *m);
float orc_monitor_max_mag(const orc_monitor_t *m);

/* ---- a7-a8: Costas sync + top-K heap ---- */
int orc_sync_score(const orc_waterfall_t *wf, const orc_candidate_t *c);
int orc_find_sync(const orc_waterfall_t *wf, int num_candidates, orc_candidate_t *heap, int min_score);

/* ---- a9-a14: per-candidate decode ---- */
void orc_extract_llr(const orc_waterfall_t *wf, const orc_candidate_t *c, float *llr174);
void orc_normalize_llr(float *llr174);
void orc_bp_decode(const float *llr174, int max_iters, uint8_t *plain174, int *errors);
uint16_t orc_crc14(const uint8_t *msg, int num_bits);
int orc_unpack77(const uint8_t *a77, char *text /* >= 35 bytes */);
int orc_decode(const orc_waterfall_t *wf, const orc_candidate_t *c, int max_iters,
               orc_message_t *msg, orc_status_t *st, float *llr_out, uint8_t *plain_out);

/* ---- a15 + whole slot: ft8_subsystem() ---- */
typedef struct {
    int n_cand;                 /* candidates returned by find_sync */
    int n_unique;               /* == *n_results of the reference    */
    /* first-seen unique messages in candidate order */
    orc_message_t msgs[512];
    float freq_hz[512];
    int score[512];
} orc_slot_report_t;
int orc_subsystem(const float *i_s, const float *q_s, int max_candidates, int max_messages, int min_score,
                  int ldpc_iters, orc_result_t *results /* max_messages, caller-zeroed */, orc_slot_report_t *rep,
                  uint8_t *wf_out /* 94208 or NULL */, orc_candidate_t *cand_out /* max_candidates or NULL */);
/* the glue alone, fed an existing waterfall (12 kHz path / stage-wise tests) */
int orc_decode_waterfall(const orc_waterfall_t *wf, int max_candidates, int max_messages, int min_score, int ldpc_iters,
                         orc_result_t *results, orc_slot_report_t *rep, orc_candidate_t *cand_out);

/* the duplicate table + CQ filter alone over decoded candidates (rtlsdr_ft8d.c:1467-1522); rep may be NULL */
int orc_spots(const orc_candidate_t *cand, const uint8_t *ok, const orc_message_t *msgs, int n_cand, int max_messages, int min_score,
              int freq_osr, orc_result_t *results /* max_messages, caller-zeroed */, orc_slot_report_t *rep);

/* ---- encoder side (input synthesis only) ---- */
int orc_pack_std(const char *call_to, const char *call_de, const char *extra, uint8_t *payload10);
void orc_pack_text(const char *text, uint8_t *payload10);
int orc_pack77(const char *msg, uint8_t *payload10); /* ref: pack77, pack.c:284-301; 0 = standard message, 1 = free text */
void orc_encode_tones(const uint8_t *payload10, uint8_t *tones79);
void orc_encode_tones_ft4(const uint8_t *payload10, uint8_t *tones105);
void orc_encode174(const uint8_t *payload10, uint8_t *bits174);

/* ---- CPU twin of the device signal synthesiser (csrc/synth.cu); layout mirrors ft8b200_signal_t ---- */
typedef struct { uint8_t payload[10]; uint8_t reserved[2]; float f0_hz, t0_sec, amp; } orc_signal_t;
void orc_synth_tones(const orc_signal_t *sig, int ft4, uint8_t *tones105);
void orc_synth_raw(const orc_signal_t *sigs, int n_sigs, float noise_lsb, uint64_t seed, int slot_index, uint8_t *iq, long long n_samples);
void orc_synth_float(int kind, int ft4, const orc_signal_t *sigs, int n_sigs, float noise_sigma, uint64_t seed, int slot_index, float *out_i,
                     float *out_q, int n_samples);

/* ---- f4: reporting formats (ft8_oracle_report.c; ref rtlsdr_ft8d.c:365-663) ---- */
int orc_pskreporter_datagram(const orc_result_t *spots, uint32_t n_spots, const char *rcall, const char *rloc, uint32_t dial_freq,
                             const char *app_version, uint32_t unixtime, uint32_t sequence, uint32_t random_id, unsigned char *out /* >= 2048 */);
void orc_webcluster_form(const orc_result_t *spot, const char *rcall, const char *rloc, uint32_t dial_freq, char *out /* 4 x 112 */);
int orc_print_spots(const orc_result_t *spots, uint32_t n_spots, uint32_t dial_freq, uint32_t unixtime, char *out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
