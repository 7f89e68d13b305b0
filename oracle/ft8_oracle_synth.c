/* oracle/ft8_oracle_synth.c -- CPU twin of the device signal synthesiser (rtlsdr-ft8d_b200/csrc/synth.cu).
 * TEST INFRASTRUCTURE ONLY (see ft8_oracle.h).  The channel symbols come from the restated ft8_encode / ft4_encode
 * (ft8_oracle_codec.c, pinned to the reference); the waveform generator is this repository's own definition -- the
 * reference's modulators (rtlsdr_ft8d.c:937-955, gen_ft8.c:28-102) use rand() and per-sample libm and are not reproducible
 * -- and is specified here in plain C: 32-bit phase accumulator per signal, 4096-entry cosine table, splitmix64 counter noise.
 * GFSK (orc_signal_t.reserved[0] = 1) follows gen_ft8.c:28-102 -- gfsk_pulse()'s own float expression, the extended end symbols, the
 * raised-cosine ramps -- with the pulse quantised to integers (q[j] = round(pulse[j] * tone-spacing word)) and the phase taken from
 * its prefix sums, so that it has a closed form per sample; tests/test_oracle_vs_ref.py compares the result with the reference's
 * synth_gfsk() (built from gen_ft8.c, oracle/_ref/libref_gen.so).
 */
#include "ft8_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define LUT_BITS 12
#define LUT_N (1 << LUT_BITS)

static float lut_f[LUT_N];
static int16_t lut_q14[LUT_N];
static int lut_ready = 0;
static void build_lut(void) {
    if (lut_ready) return;
    for (int i = 0; i < LUT_N; ++i) {
        const double c = cos(2.0 * M_PI * (double)i / (double)LUT_N);
        lut_f[i] = (float)c;
        lut_q14[i] = (int16_t)lround(16384.0 * c);
    }
    lut_ready = 1;
}

static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static uint64_t slot_key(uint64_t seed, int slot) { return splitmix64(seed + 0x632BE59BD9B4E019ull * (uint64_t)(slot + 1)); }
static int byte_sum4(uint32_t w) { return (int)(w & 255u) + (int)((w >> 8) & 255u) + (int)((w >> 16) & 255u) + (int)(w >> 24); }

typedef struct {
    long long s0;
    uint32_t fw[8];
    uint32_t pstart[105];
    float amp;
    int32_t amp_q8;
    int n_sym;
    int gfsk;
    uint8_t tones[105];
} sig_t;

/* integer GFSK pulse of one symbol length: prefix sums P[0..3L] and the ramp of L/8 samples (ref: gfsk_pulse, gen_ft8.c:28-38; :96-101) */
typedef struct { int L; float bt; uint32_t step; uint32_t *P; float *env_f; int32_t *env_q; } gfsk_t;
static gfsk_t g_gfsk[4];
static const gfsk_t *gfsk_table(int L, float bt, uint32_t step) {
    gfsk_t *t = NULL;
    for (int k = 0; k < 4; ++k) {
        if (g_gfsk[k].P && g_gfsk[k].L == L && g_gfsk[k].bt == bt && g_gfsk[k].step == step) return &g_gfsk[k];
        if (!g_gfsk[k].P && !t) t = &g_gfsk[k];
    }
    if (!t) { t = &g_gfsk[0]; free(t->P); free(t->env_f); free(t->env_q); }
    t->L = L; t->bt = bt; t->step = step;
    t->P = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)3 * L + 1));
    uint32_t acc = 0;
    for (int j = 0; j < 3 * L; ++j) {
        const float tt = j / (float)L - 1.5f;
        const float arg1 = 5.336446f * bt * (tt + 0.5f), arg2 = 5.336446f * bt * (tt - 0.5f);
        const float pulse = (erff(arg1) - erff(arg2)) / 2;
        t->P[j] = acc;
        acc += (uint32_t)llround((double)pulse * (double)step);
    }
    t->P[3 * L] = acc;
    const int n_ramp = L / 8;
    t->env_f = (float *)malloc(sizeof(float) * (size_t)(n_ramp > 0 ? n_ramp : 1));
    t->env_q = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_ramp > 0 ? n_ramp : 1));
    for (int i = 0; i < n_ramp; ++i) {
        t->env_f[i] = (1 - cosf(2 * (float)M_PI * i / (2 * n_ramp))) / 2;
        t->env_q[i] = (int32_t)lround((double)t->env_f[i] * 32768.0);
    }
    return t;
}

static const gfsk_t *prepare(const orc_signal_t *in, sig_t *s, double fs, int sym_len, double tone_hz, double f_shift, int ft4) {
    memset(s, 0, sizeof(*s));
    s->gfsk = in->reserved[0] != 0;
    s->s0 = llround((double)in->t0_sec * fs);
    for (int t = 0; t < 8; ++t) s->fw[t] = (uint32_t)(int64_t)llround(((double)in->f0_hz + t * tone_hz + f_shift) / fs * 4294967296.0);
    s->amp = in->amp;
    s->amp_q8 = (int32_t)lround((double)in->amp * 256.0);
    if (ft4) { orc_encode_tones_ft4(in->payload, s->tones); s->n_sym = 105; }
    else { orc_encode_tones(in->payload, s->tones); s->n_sym = 79; }
    const gfsk_t *gf = s->gfsk ? gfsk_table(sym_len, ft4 ? 1.0f : 2.0f, (uint32_t)llround(tone_hz / fs * 4294967296.0)) : NULL;
    uint32_t ph = 0;
    for (int i = 0; i < s->n_sym; ++i) {
        s->pstart[i] = ph;
        if (gf) {
            const uint32_t tp = s->tones[i > 0 ? i - 1 : 0], tc = s->tones[i], tn = s->tones[i + 1 < s->n_sym ? i + 1 : i];
            ph += (uint32_t)sym_len * s->fw[0] + tn * gf->P[sym_len] + tc * (gf->P[2 * sym_len] - gf->P[sym_len]) + tp * (gf->P[3 * sym_len] - gf->P[2 * sym_len]);
        } else {
            ph += (uint32_t)sym_len * s->fw[s->tones[i]];
        }
    }
    return gf;
}

/* phase of the signal at sample n (0 = silent there); *ramp = index into the GFSK envelope table or -1 */
static int phase_at(const sig_t *s, const gfsk_t *gf, long long n, int sym_len, uint32_t *ph, int *ramp) {
    const long long rel = n - s->s0, total = (long long)s->n_sym * sym_len;
    if (rel < 0 || rel >= total) return 0;
    const int k = (int)(rel / sym_len);
    const uint32_t j = (uint32_t)(rel - (long long)k * sym_len);
    *ramp = -1;
    if (s->gfsk) {
        const uint32_t tp = s->tones[k > 0 ? k - 1 : 0], tc = s->tones[k], tn = s->tones[k + 1 < s->n_sym ? k + 1 : k];
        *ph = s->pstart[k] + j * s->fw[0] + tn * gf->P[j] + tc * (gf->P[j + sym_len] - gf->P[sym_len]) + tp * (gf->P[j + 2 * sym_len] - gf->P[2 * sym_len]);
        const int n_ramp = sym_len / 8;
        if (rel < n_ramp) *ramp = (int)rel;
        else if (rel >= total - n_ramp) *ramp = (int)(total - 1 - rel);
    } else {
        *ph = s->pstart[k] + j * s->fw[s->tones[k]];
    }
    return 1;
}

void orc_synth_tones(const orc_signal_t *sig, int ft4, uint8_t *tones105) {
    sig_t s;
    prepare(sig, &s, 12000.0, ft4 ? 576 : 1920, ft4 ? 1.0 / 0.048 : 6.25, 0.0, ft4);
    memcpy(tones105, s.tones, 105);
}

void orc_synth_raw(const orc_signal_t *sigs, int n_sigs, float noise_lsb, uint64_t seed, int slot_index, uint8_t *iq, long long n_samples) {
    build_lut();
    sig_t *s = (sig_t *)calloc((size_t)(n_sigs > 0 ? n_sigs : 1), sizeof(sig_t));
    const gfsk_t *gf = NULL;
    for (int g = 0; g < n_sigs; ++g) { const gfsk_t *t = prepare(&sigs[g], &s[g], 2400000.0, 384000, 6.25, -600000.0, 0); if (t) gf = t; }
    const int noise_q8 = (int)lround((double)noise_lsb * 65536.0 / 147.79715829474123);
    const uint64_t key = slot_key(seed, slot_index);
    for (long long n = 0; n < n_samples; ++n) {
        int vi = 0, vq = 0;
        for (int g = 0; g < n_sigs; ++g) {
            uint32_t ph;
            int ramp;
            if (!phase_at(&s[g], gf, n, 384000, &ph, &ramp)) continue;
            const int idx = (int)(ph >> (32 - LUT_BITS));
            int amp = s[g].amp_q8;
            if (ramp >= 0) amp = (int)(((long long)amp * gf->env_q[ramp] + 16384) >> 15);
            vi += (amp * (int)lut_q14[idx] + (1 << 21)) >> 22;
            vq += (amp * (int)lut_q14[(idx - LUT_N / 4) & (LUT_N - 1)] + (1 << 21)) >> 22;
        }
        const uint64_t h = splitmix64(key + (uint64_t)n);
        const int ni = ((byte_sum4((uint32_t)h) - 510) * noise_q8 + (1 << 15)) >> 16;
        const int nq = ((byte_sum4((uint32_t)(h >> 32)) - 510) * noise_q8 + (1 << 15)) >> 16;
        int bi = 128 + vi + ni, bq = 128 + vq + nq;
        iq[2 * n] = (uint8_t)(bi < 0 ? 0 : (bi > 255 ? 255 : bi));
        iq[2 * n + 1] = (uint8_t)(bq < 0 ? 0 : (bq > 255 ? 255 : bq));
    }
    free(s);
}

/* kind 1: complex baseband at 3200 sps (out_q != NULL); kind 2: real audio at 12 kHz, FT8 or FT4 */
void orc_synth_float(int kind, int ft4, const orc_signal_t *sigs, int n_sigs, float noise_sigma, uint64_t seed, int slot_index, float *out_i,
                     float *out_q, int n_samples) {
    build_lut();
    const int sym_len = kind == 1 ? 512 : (ft4 ? 576 : 1920);
    sig_t *s = (sig_t *)calloc((size_t)(n_sigs > 0 ? n_sigs : 1), sizeof(sig_t));
    const gfsk_t *gf = NULL;
    for (int g = 0; g < n_sigs; ++g) {
        const gfsk_t *t = prepare(&sigs[g], &s[g], kind == 1 ? 3200.0 : 12000.0, sym_len, (kind == 2 && ft4) ? 1.0 / 0.048 : 6.25, 0.0, kind == 2 && ft4);
        if (t) gf = t;
    }
    const float scale = (float)((double)noise_sigma / 209.02153956946134);
    const uint64_t key = slot_key(seed, slot_index);
    for (int n = 0; n < n_samples; ++n) {
        const uint64_t h0 = splitmix64(key + 2ull * (uint64_t)n);
        float vi = (float)(byte_sum4((uint32_t)h0) + byte_sum4((uint32_t)(h0 >> 32)) - 1020) * scale;
        float vq = 0.0f;
        if (kind == 1) {
            const uint64_t h1 = splitmix64(key + 2ull * (uint64_t)n + 1ull);
            vq = (float)(byte_sum4((uint32_t)h1) + byte_sum4((uint32_t)(h1 >> 32)) - 1020) * scale;
        }
        for (int g = 0; g < n_sigs; ++g) {
            uint32_t ph;
            int ramp;
            if (!phase_at(&s[g], gf, (long long)n, sym_len, &ph, &ramp)) continue;
            const int idx = (int)(ph >> (32 - LUT_BITS));
            const float amp = ramp >= 0 ? s[g].amp * gf->env_f[ramp] : s[g].amp;
            vi = vi + amp * lut_f[idx];
            if (kind == 1) vq = vq + amp * lut_f[(idx - LUT_N / 4) & (LUT_N - 1)];
        }
        out_i[n] = vi;
        if (kind == 1) out_q[n] = vq;
    }
    free(s);
}
