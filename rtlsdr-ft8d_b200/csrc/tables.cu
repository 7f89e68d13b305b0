// tables.cu -- host-side construction of the constant tables the kernels consume.
// The reference computes these with the host libm (sinf, cos, sin, log10f); doing the same here,
// once per context, is what makes the device results bit-identical to it.
#include "common.cuh"

#include <math.h>
#include <string.h>

namespace ft8b200 {

// CIC compensation FIR, R=750 M=1 N=2 F0=0.92 L=54; symmetric, centre tap 0.5 (filter design data).
// ref: rtlsdr_ft8d.c:93-110 -- the reference writes double literals that are converted to float.
void build_fir(float *z) {
    static const double half[28] = {
        -0.0025719973, 0.0010118403,  0.0009110571,  -0.0034940765, 0.0069713409,  -0.0114242790, 0.0167023466,
        -0.0223683056, 0.0276808966,  -0.0316243672, 0.0329894230,  -0.0305042011, 0.0230074504,  -0.0096499429,
        -0.0098950502, 0.0352349632,  -0.0650990428, 0.0972406918,  -0.1284211497, 0.1544893973,  -0.1705667465,
        0.1713383321,  -0.1514501610, 0.1060148823,  -0.0312560926, -0.0745846391, 0.2096088743,  -0.3638689868,
    };
    for (int j = 0; j < 28; ++j) z[j] = z[56 - j] = (float)half[j];
    z[28] = 0.5f;
}

// ref: initFFTW(), rtlsdr_ft8d.c:331-334 ("hann" there is a half-sine)
void build_window1024(float *w) {
    for (int i = 0; i < kNfft; ++i) w[i] = sinf((float)((M_PI / kNfft) * i));
}

// ref: kiss_fft_alloc(), ft8_lib/fft/kiss_fft.c:351-357
void build_twiddles(int n, float2 *tw) {
    for (int k = 0; k < n; ++k) {
        const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
        const double phase = -2 * pi * k / n;
        tw[k].x = (float)cos(phase);
        tw[k].y = (float)sin(phase);
    }
}

// The reference quantises x -> clamp((int)(2*(10*log10f(x)) + 240), 0, 255) (rtlsdr_ft8d.c:1416,1425-1427;
// decode_ft8.c:203-208).  That map is a monotone step function of x, so it is fully described by the
// 255 float thresholds where it steps; they are found here by bisection over float bit patterns with the
// host's own log10f, and the kernels only compare against them (no device log needed, no libm mismatch).
static inline int quantise_ref(float x) {
    const float db = 10.0f * log10f(x);
    const int scaled = (int)(2 * db + 240);
    return scaled < 0 ? 0 : (scaled > 255 ? 255 : scaled);
}
void build_db_thresholds(float *t) {
    t[0] = 0.0f;
    for (int k = 1; k <= 255; ++k) {
        float lo = 1e-13f, hi = 1e30f;  // quantise to 0 and 255
        uint32_t ulo, uhi;
        memcpy(&ulo, &lo, 4);
        memcpy(&uhi, &hi, 4);
        while (uhi - ulo > 1) {
            const uint32_t um = ulo + (uhi - ulo) / 2;
            float m;
            memcpy(&m, &um, 4);
            if (quantise_ref(m) >= k) uhi = um; else ulo = um;
        }
        memcpy(&t[k], &uhi, 4);
    }
    t[256] = INFINITY;
}

}  // namespace ft8b200
