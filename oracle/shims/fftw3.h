/* Stand-in for <fftw3.h>: maps the six single-precision FFTW calls the reference
 * daemon makes (rtlsdr_ft8d.c:314-347, :1411) onto the complex kiss_fft that the
 * reference vendors in ft8_lib/fft/.  FFTW3f itself is an external, unpinned distro
 * package whose plan (hence rounding) is machine dependent; kiss_fft is the only FFT
 * whose arithmetic ships with the reference, so it is the reproducible oracle
 * ("FFT rounding: parity unpinned" w.r.t. FFTW proper). */
#ifndef ORACLE_SHIM_FFTW3_H
#define ORACLE_SHIM_FFTW3_H
#include <stdio.h>
#include <stdlib.h>
#include "ft8_lib/fft/kiss_fft.h"
typedef float fftwf_complex[2];
typedef struct oracle_fftw_plan { kiss_fft_cfg cfg; fftwf_complex *in, *out; } *fftwf_plan;
#define FFTW_FORWARD (-1)
#define FFTW_ESTIMATE (1U << 6)
static inline void *fftwf_malloc(size_t n) { return malloc(n); }
static inline void fftwf_free(void *p) { free(p); }
static inline int fftwf_import_wisdom_from_file(FILE *f) { (void)f; return 0; }
static inline void fftwf_export_wisdom_to_file(FILE *f) { (void)f; }
static inline fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags) {
    (void)flags;
    fftwf_plan p = (fftwf_plan)malloc(sizeof(*p));
    p->cfg = kiss_fft_alloc(n, sign == FFTW_FORWARD ? 0 : 1, NULL, NULL);
    p->in = in; p->out = out;
    return p;
}
static inline void fftwf_execute(const fftwf_plan p) {
    kiss_fft(p->cfg, (const kiss_fft_cpx *)p->in, (kiss_fft_cpx *)p->out);
}
static inline void fftwf_destroy_plan(fftwf_plan p) { if (p) { free(p->cfg); free(p); } }
#endif
