/* oracle/ft8_oracle.c -- CPU restatement of the rtlsdr-ft8d hot path (see ft8_oracle.h).
 * TEST INFRASTRUCTURE ONLY: never linked into or called by the product library.
 * Build: gcc -O3 -std=gnu17 -ffp-contract=off -fwrapv  (oracle/Makefile).
 * "ref:" comments give the reference file:line each block follows.
 */
#include "ft8_oracle.h"
#include "ft8_tables.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================================
 * a1-a3  decimator                                             ref: rtlsdr_ft8d.c:76-202
 * ====================================================================================== */

/* CIC compensation FIR, R=750 M=1 N=2 F0=0.92 L=54 (filter design data). Symmetric about
 * tap 28 (= 0.5).  ref: rtlsdr_ft8d.c:93-110 (written there as double literals -> float). */
static const float kFirHalf[28] = {
    -0.0025719973, 0.0010118403,  0.0009110571,  -0.0034940765, 0.0069713409,  -0.0114242790, 0.0167023466,
    -0.0223683056, 0.0276808966,  -0.0316243672, 0.0329894230,  -0.0305042011, 0.0230074504,  -0.0096499429,
    -0.0098950502, 0.0352349632,  -0.0650990428, 0.0972406918,  -0.1284211497, 0.1544893973,  -0.1705667465,
    0.1713383321,  -0.1514501610, 0.1060148823,  -0.0312560926, -0.0745846391, 0.2096088743,  -0.3638689868,
};
static float g_fir[ORC_FIR_TAPS];
static int g_fir_ready = 0;
const float *orc_fir_coefs(void) {
    if (!g_fir_ready) {
        for (int j = 0; j < 28; ++j) { g_fir[j] = kFirHalf[j]; g_fir[56 - j] = kFirHalf[j]; }
        g_fir[28] = 0.5f;
        g_fir_ready = 1;
    }
    return g_fir;
}

void orc_decim_reset(orc_decim_t *st) { memset(st, 0, sizeof(*st)); }

/* int8 negate as the reference's `sigIn[k] = -tmp` store does it: two's-complement wrap,
 * so -(-128) stays -128.  ref: rtlsdr_ft8d.c:132-139 */
static inline int8_t neg_wrap8(int8_t v) { return (int8_t)(uint8_t)(0u - (uint8_t)v); }

void orc_decim_feed(orc_decim_t *st, const uint8_t *iq, size_t nbytes, float *i_out, float *q_out, int32_t *y2i,
                    int32_t *y2q, size_t cap, size_t *count) {
    const float *z = orc_fir_coefs();
    const size_t nsamp = nbytes / 2;
    for (size_t n = 0; n < nsamp; ++n) {
        /* fs/4 mixer: sample n of THIS call is multiplied by j^n.  ref: :128-140 */
        const int8_t a = (int8_t)(iq[2 * n] ^ 0x80), b = (int8_t)(iq[2 * n + 1] ^ 0x80);
        int8_t mi, mq;
        switch (n & 3u) {
        case 0: mi = a; mq = b; break;
        case 1: mi = neg_wrap8(b); mq = a; break;
        case 2: mi = neg_wrap8(a); mq = neg_wrap8(b); break;
        default: mi = b; mq = neg_wrap8(a); break;
        }
        /* two integrators per rail, int32 with wrap.  ref: :150-153 */
        st->ix1 = (int32_t)((uint32_t)st->ix1 + (uint32_t)(int32_t)mi);
        st->qx1 = (int32_t)((uint32_t)st->qx1 + (uint32_t)(int32_t)mq);
        st->ix2 = (int32_t)((uint32_t)st->ix2 + (uint32_t)st->ix1);
        st->qx2 = (int32_t)((uint32_t)st->qx2 + (uint32_t)st->qx1);
        /* keep one sample in 751.  ref: :156-160 */
        if (++st->decim_index <= 750u) continue;
        st->decim_index = 0;
        /* two combs, differential delay 2.  ref: :163-176 */
        const int32_t iy1 = (int32_t)((uint32_t)st->ix2 - (uint32_t)st->it1z);
        st->it1z = st->it1y; st->it1y = st->ix2;
        const int32_t qy1 = (int32_t)((uint32_t)st->qx2 - (uint32_t)st->qt1z);
        st->qt1z = st->qt1y; st->qt1y = st->qx2;
        const int32_t iy2 = (int32_t)((uint32_t)iy1 - (uint32_t)st->it2z);
        st->it2z = st->it2y; st->it2y = iy1;
        const int32_t qy2 = (int32_t)((uint32_t)qy1 - (uint32_t)st->qt2z);
        st->qt2z = st->qt2y; st->qt2y = qy1;
        /* 57-tap FIR, strictly sequential float MACs (no FMA).  ref: :179-192 */
        float si = 0.0f, sq = 0.0f;
        for (int j = 0; j < 56; ++j) {
            si += st->fir_i[j] * z[j];
            sq += st->fir_q[j] * z[j];
        }
        memmove(st->fir_i, st->fir_i + 1, 55 * sizeof(float));
        memmove(st->fir_q, st->fir_q + 1, 55 * sizeof(float));
        st->fir_i[55] = (float)iy2;
        st->fir_q[55] = (float)qy2;
        si += st->fir_i[55] * z[56];
        sq += st->fir_q[55] * z[56];
        /* scale in double, store as float while there is room.  ref: :195-200 */
        if (*count < cap) {
            i_out[*count] = (float)((double)si / (32768.0 * 750));
            q_out[*count] = (float)((double)sq / (32768.0 * 750));
            if (y2i) y2i[*count] = iy2;
            if (y2q) y2q[*count] = qy2;
            ++*count;
        }
        ++st->n_out;
    }
}

/* ======================================================================================
 * a4  slot conditioning                                       ref: rtlsdr_ft8d.c:242-263
 * ====================================================================================== */
float orc_condition(float *i_s, float *q_s, size_t n_valid, size_t n_total) {
    for (size_t k = n_valid; k < n_total; ++k) { i_s[k] = 0.0f; q_s[k] = 0.0f; }
    float peak = 1e-24f;
    for (size_t k = 0; k < n_total; ++k) {
        const float ai = (float)fabs(i_s[k]), aq = (float)fabs(q_s[k]);
        if (ai > peak) peak = ai;
        if (aq > peak) peak = aq;
    }
    const float scale = (float)(0.5 / (double)peak);
    for (size_t k = 0; k < n_total; ++k) { i_s[k] *= scale; q_s[k] *= scale; }
    return scale;
}

/* ======================================================================================
 * FFT: kiss_fft's float arithmetic restated (decimation in time, radix 4 first, then
 * 2, 3, 5, then the generic butterfly), same twiddle table, same operation order inside every butterfly.
 * ref: ft8_lib/fft/kiss_fft.c:15-382, _kiss_fft_guts.h:81-83 (C_MUL), kiss_fftr.c:22-115
 * ====================================================================================== */
typedef struct { float r, i; } cpx;

typedef struct {
    int n;
    int nfac;
    int radix[32], rem[32];
    cpx *tw;
} fft_plan_t;

static fft_plan_t *plan_make(int n) {
    fft_plan_t *p = (fft_plan_t *)calloc(1, sizeof(*p));
    p->n = n;
    p->tw = (cpx *)malloc(sizeof(cpx) * (size_t)n);
    for (int k = 0; k < n; ++k) { /* ref: kiss_fft.c:351-357 */
        const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
        const double phase = -2 * pi * k / n;
        p->tw[k].r = (float)cos(phase);
        p->tw[k].i = (float)sin(phase);
    }
    /* ref: kf_factor, kiss_fft.c:303-324 */
    int left = n, q = 4;
    const double lim = floor(sqrt((double)n));
    do {
        while (left % q) {
            q = (q == 4) ? 2 : (q == 2) ? 3 : q + 2;
            if (q > lim) q = left;
        }
        left /= q;
        p->radix[p->nfac] = q;
        p->rem[p->nfac] = left;
        ++p->nfac;
    } while (left > 1);
    return p;
}
static void plan_free(fft_plan_t *p) { if (p) { free(p->tw); free(p); } }

static inline cpx cmul(cpx a, cpx b) { /* ref: C_MUL */
    cpx m;
    m.r = a.r * b.r - a.i * b.i;
    m.i = a.r * b.i + a.i * b.r;
    return m;
}
static inline cpx cadd(cpx a, cpx b) { cpx m = { a.r + b.r, a.i + b.i }; return m; }
static inline cpx csub(cpx a, cpx b) { cpx m = { a.r - b.r, a.i - b.i }; return m; }

static void combine(const fft_plan_t *p, cpx *F, int stride, int radix, int m) {
    const cpx *tw = p->tw;
    if (radix == 4) { /* ref: kf_bfly4, kiss_fft.c:38-84 (forward) */
        for (int k = 0; k < m; ++k) {
            cpx *f0 = F + k, *f1 = f0 + m, *f2 = f0 + 2 * m, *f3 = f0 + 3 * m;
            const cpx a = cmul(*f1, tw[k * stride]);
            const cpx b = cmul(*f2, tw[2 * k * stride]);
            const cpx c = cmul(*f3, tw[3 * k * stride]);
            const cpx d5 = csub(*f0, b);
            *f0 = cadd(*f0, b);
            const cpx s3 = cadd(a, c);
            const cpx s4 = csub(a, c);
            *f2 = csub(*f0, s3);
            *f0 = cadd(*f0, s3);
            f1->r = d5.r + s4.i; f1->i = d5.i - s4.r;
            f3->r = d5.r - s4.i; f3->i = d5.i + s4.r;
        }
    } else if (radix == 2) { /* ref: kf_bfly2, kiss_fft.c:15-36 */
        for (int k = 0; k < m; ++k) {
            cpx *f0 = F + k, *f1 = f0 + m;
            const cpx t = cmul(*f1, tw[k * stride]);
            *f1 = csub(*f0, t);
            *f0 = cadd(*f0, t);
        }
    } else if (radix == 3) { /* ref: kf_bfly3, kiss_fft.c:86-128 */
        const cpx e3 = tw[stride * m];
        for (int k = 0; k < m; ++k) {
            cpx *f0 = F + k, *f1 = f0 + m, *f2 = f0 + 2 * m;
            const cpx s1 = cmul(*f1, tw[k * stride]);
            const cpx s2 = cmul(*f2, tw[2 * k * stride]);
            const cpx s3 = cadd(s1, s2);
            cpx s0 = csub(s1, s2);
            f1->r = (float)(f0->r - s3.r * .5);
            f1->i = (float)(f0->i - s3.i * .5);
            s0.r *= e3.i; s0.i *= e3.i;
            *f0 = cadd(*f0, s3);
            f2->r = f1->r + s0.i;
            f2->i = f1->i - s0.r;
            f1->r -= s0.i;
            f1->i += s0.r;
        }
    } else if (radix == 5) { /* ref: kf_bfly5, kiss_fft.c:130-190 */
        const cpx ya = tw[stride * m], yb = tw[stride * 2 * m];
        for (int u = 0; u < m; ++u) {
            cpx *f0 = F + u, *f1 = f0 + m, *f2 = f0 + 2 * m, *f3 = f0 + 3 * m, *f4 = f0 + 4 * m;
            const cpx s0 = *f0;
            const cpx s1 = cmul(*f1, tw[u * stride]);
            const cpx s2 = cmul(*f2, tw[2 * u * stride]);
            const cpx s3 = cmul(*f3, tw[3 * u * stride]);
            const cpx s4 = cmul(*f4, tw[4 * u * stride]);
            const cpx s7 = cadd(s1, s4), s10 = csub(s1, s4), s8 = cadd(s2, s3), s9 = csub(s2, s3);
            f0->r += s7.r + s8.r;
            f0->i += s7.i + s8.i;
            cpx s5, s6, s11, s12;
            s5.r = s0.r + s7.r * ya.r + s8.r * yb.r;
            s5.i = s0.i + s7.i * ya.r + s8.i * yb.r;
            s6.r = s10.i * ya.i + s9.i * yb.i;
            s6.i = -(s10.r * ya.i) - s9.r * yb.i;
            *f1 = csub(s5, s6);
            *f4 = cadd(s5, s6);
            s11.r = s0.r + s7.r * yb.r + s8.r * ya.r;
            s11.i = s0.i + s7.i * yb.r + s8.i * ya.r;
            s12.r = -(s10.i * yb.i) + s9.i * ya.i;
            s12.i = s10.r * yb.i - s9.r * ya.i;
            *f2 = cadd(s11, s12);
            *f3 = csub(s11, s12);
        }
    } else { /* ref: kf_bfly_generic, kiss_fft.c:192-229 -- any other radix (7, 11, 13 ... or what is left of n when it is prime) */
        const int Norig = p->n;
        cpx *scratch = (cpx *)malloc(sizeof(cpx) * (size_t)radix);
        for (int u = 0; u < m; ++u) {
            int k = u;
            for (int q1 = 0; q1 < radix; ++q1) { scratch[q1] = F[k]; k += m; }
            k = u;
            for (int q1 = 0; q1 < radix; ++q1) {
                int twidx = 0;
                F[k] = scratch[0];
                for (int q = 1; q < radix; ++q) {
                    twidx += stride * k;
                    if (twidx >= Norig) twidx -= Norig;
                    const cpx t = cmul(scratch[q], tw[twidx]);
                    F[k].r += t.r; /* C_ADDTO */
                    F[k].i += t.i;
                }
                k += m;
            }
        }
        free(scratch);
    }
}

/* ref: kf_work, kiss_fft.c:241-296 (in_stride == 1) */
static void fft_rec(const fft_plan_t *p, cpx *out, const cpx *in, int stride, int level) {
    const int radix = p->radix[level], m = p->rem[level];
    if (m == 1) {
        for (int k = 0; k < radix; ++k) out[k] = in[(size_t)k * stride];
    } else {
        for (int k = 0; k < radix; ++k) fft_rec(p, out + (size_t)k * m, in + (size_t)k * stride, stride * radix, level + 1);
    }
    combine(p, out, stride, radix, m);
}

static fft_plan_t *g_plans[8];
static fft_plan_t *plan_get(int n) {
    for (int k = 0; k < 8; ++k) {
        if (g_plans[k] && g_plans[k]->n == n) return g_plans[k];
        if (!g_plans[k]) return g_plans[k] = plan_make(n);
    }
    plan_free(g_plans[0]);
    return g_plans[0] = plan_make(n);
}

void orc_fft_c2c(int n, const float *in_ri, float *out_ri) {
    fft_rec(plan_get(n), (cpx *)out_ri, (const cpx *)in_ri, 1, 0);
}

/* real-input FFT through an n/2-point complex FFT.  ref: kiss_fftr.c:22-115 */
void orc_fft_r2c(int n, const float *in, float *out_ri) {
    const int h = n / 2;
    cpx *tmp = (cpx *)malloc(sizeof(cpx) * (size_t)h);
    cpx *st = (cpx *)malloc(sizeof(cpx) * (size_t)(h / 2 + 1));
    cpx *out = (cpx *)out_ri;
    for (int k = 0; k < h / 2; ++k) { /* ref: kiss_fftr.c:50-56 */
        const double phase = -3.14159265358979323846264338327 * ((double)(k + 1) / h + .5);
        st[k].r = (float)cos(phase);
        st[k].i = (float)sin(phase);
    }
    fft_rec(plan_get(h), tmp, (const cpx *)in, 1, 0);
    out[0].r = tmp[0].r + tmp[0].i;
    out[h].r = tmp[0].r - tmp[0].i;
    out[0].i = out[h].i = 0;
    for (int k = 1; k <= h / 2; ++k) { /* ref: kiss_fftr.c:97-113 */
        const cpx fpk = tmp[k];
        cpx fpnk = { tmp[h - k].r, -tmp[h - k].i };
        const cpx f1 = cadd(fpk, fpnk), f2 = csub(fpk, fpnk);
        const cpx t = cmul(f2, st[k - 1]);
        out[k].r = (float)((f1.r + t.r) * .5);
        out[k].i = (float)((f1.i + t.i) * .5);
        out[h - k].r = (float)((f1.r - t.r) * .5);
        out[h - k].i = (float)((t.i - f1.i) * .5);
    }
    free(tmp);
    free(st);
}

/* ======================================================================================
 * a5  daemon waterfall                              ref: rtlsdr_ft8d.c:314-335, 1395-1435
 * ====================================================================================== */
void orc_sine_window(float *w, int n) { /* "hann" in the reference, actually a half sine */
    for (int k = 0; k < n; ++k) w[k] = sinf((float)((M_PI / n) * k));
}

/* dB -> u8: 0.5 dB steps, 0 == -120 dB.  ref: rtlsdr_ft8d.c:1416,1425-1427; decode_ft8.c:203-208 */
uint8_t orc_quantize_db(float x) {
    const float db = 10.0f * log10f(x);
    const int scaled = (int)(2 * db + 240);
    return (uint8_t)(scaled < 0 ? 0 : (scaled > 255 ? 255 : scaled));
}

void orc_db_thresholds(float *t) {
    union { float f; uint32_t u; } lo, hi, mid;
    t[0] = 0.0f;
    for (int k = 1; k <= 255; ++k) {
        lo.f = 1e-13f; /* quantizes to 0 */
        hi.f = 1e30f;  /* quantizes to 255 */
        while (hi.u - lo.u > 1) {
            mid.u = lo.u + (hi.u - lo.u) / 2;
            if (orc_quantize_db(mid.f) >= k) hi = mid; else lo = mid;
        }
        t[k] = hi.f;
    }
    t[256] = INFINITY;
}

void orc_waterfall_daemon(const float *i_s, const float *q_s, uint8_t *mag) {
    static float win[1024];
    static int win_ready = 0;
    if (!win_ready) { orc_sine_window(win, 1024); win_ready = 1; }
    cpx in[1024], out[1024];
    size_t o = 0;
    for (int blk = 0; blk < 92; ++blk) {
        for (int ts = 0; ts < 2; ++ts) {
            const int start = blk * 512 + ts * 256;
            for (int k = 0; k < 1024; ++k) {
                in[k].r = i_s[start + k] * win[k];
                in[k].i = q_s[start + k] * win[k];
            }
            orc_fft_c2c(1024, (const float *)in, (float *)out);
            for (int fs = 0; fs < 2; ++fs) {
                for (int b = 0; b < 256; ++b) {
                    const cpx v = out[b * 2 + fs];
                    const float mag2 = v.r * v.r + v.i * v.i;
                    mag[o++] = orc_quantize_db(1E-12f + mag2 * 4.0f / (float)(1024u * 1024u));
                }
            }
        }
    }
}

/* ======================================================================================
 * a5'  ft8_lib monitor (12 kHz real audio)                  ref: decode_ft8.c:35-39, 63-224
 * ====================================================================================== */
struct orc_monitor {
    int block_size, subblock_size, nfft;
    float fft_norm, max_mag;
    float *window, *last_frame;
    orc_waterfall_t wf;
};

orc_monitor_t *orc_monitor_new(int sample_rate, int time_osr, int freq_osr, int protocol) {
    orc_monitor_t *m = (orc_monitor_t *)calloc(1, sizeof(*m));
    const float slot_time = (protocol == 0) ? 7.5f : 15.0f;
    const float symbol_period = (protocol == 0) ? 0.048f : 0.160f;
    m->block_size = (int)(sample_rate * symbol_period);
    m->subblock_size = m->block_size / time_osr;
    m->nfft = m->block_size * freq_osr;
    m->fft_norm = 2.0f / m->nfft;
    m->window = (float *)malloc(sizeof(float) * (size_t)m->nfft);
    for (int k = 0; k < m->nfft; ++k) { /* Hann = sin^2.  ref: decode_ft8.c:35-39 */
        const float x = sinf((float)M_PI * k / m->nfft);
        m->window[k] = x * x;
    }
    /* the reference mallocs this without clearing it (decode_ft8.c:131); zero is the only
     * deterministic choice and is what oracle/ref_mon_harness.c forces on the reference too */
    m->last_frame = (float *)calloc((size_t)m->nfft, sizeof(float));
    m->wf.max_blocks = (int)(slot_time / symbol_period);
    m->wf.num_blocks = 0;
    m->wf.num_bins = (int)(sample_rate * symbol_period / 2);
    m->wf.time_osr = time_osr;
    m->wf.freq_osr = freq_osr;
    m->wf.block_stride = time_osr * freq_osr * m->wf.num_bins;
    m->wf.mag = (uint8_t *)calloc((size_t)m->wf.max_blocks * (size_t)m->wf.block_stride, 1);
    m->wf.protocol = protocol;
    m->max_mag = -120.0f;
    return m;
}
void orc_monitor_free(orc_monitor_t *m) {
    if (!m) return;
    free(m->wf.mag); free(m->window); free(m->last_frame); free(m);
}
void orc_monitor_reset(orc_monitor_t *m) { m->wf.num_blocks = 0; m->max_mag = 0; }
void orc_monitor_info(const orc_monitor_t *m, int *o) {
    o[0] = m->block_size; o[1] = m->subblock_size; o[2] = m->nfft; o[3] = m->wf.max_blocks; o[4] = m->wf.num_blocks;
    o[5] = m->wf.num_bins; o[6] = m->wf.time_osr; o[7] = m->wf.freq_osr; o[8] = m->wf.block_stride;
}
const uint8_t *orc_monitor_mag(const orc_monitor_t *m) { return m->wf.mag; }
float orc_monitor_max_mag(const orc_monitor_t *m) { return m->max_mag; }

void orc_monitor_process(orc_monitor_t *m, const float *frame) { /* ref: decode_ft8.c:162-218 */
    if (m->wf.num_blocks >= m->wf.max_blocks) return;
    const int n = m->nfft, hop = m->subblock_size;
    float *td = (float *)malloc(sizeof(float) * (size_t)n);
    cpx *fd = (cpx *)malloc(sizeof(cpx) * (size_t)(n / 2 + 1));
    size_t o = (size_t)m->wf.num_blocks * (size_t)m->wf.block_stride;
    int fp = 0;
    for (int ts = 0; ts < m->wf.time_osr; ++ts) {
        memmove(m->last_frame, m->last_frame + hop, sizeof(float) * (size_t)(n - hop));
        for (int k = n - hop; k < n; ++k) m->last_frame[k] = frame[fp++];
        for (int k = 0; k < n; ++k) td[k] = m->fft_norm * m->window[k] * m->last_frame[k];
        orc_fft_r2c(n, td, (float *)fd);
        for (int fs = 0; fs < m->wf.freq_osr; ++fs) {
            for (int b = 0; b < m->wf.num_bins; ++b) {
                const cpx v = fd[b * m->wf.freq_osr + fs];
                const float mag2 = v.i * v.i + v.r * v.r;
                const float x = 1E-12f + mag2;
                m->wf.mag[o++] = orc_quantize_db(x);
                const float db = 10.0f * log10f(x);
                if (db > m->max_mag) m->max_mag = db;
            }
        }
    }
    ++m->wf.num_blocks;
    free(td);
    free(fd);
}
