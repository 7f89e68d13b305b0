// monitor.cu -- ft8_lib's "monitor": real audio (12 kHz) -> Hann-windowed STFT -> uint8 dB waterfall.
// Replaces monitor_init/process/reset/free and the kiss_fftr loop, /root/reference/ft8_lib/decode_ft8.c:35-39,
// 111-224, ft8_lib/fft/kiss_fftr.c:22-115, ft8_lib/fft/kiss_fft.c:15-382.
//
// A frame of nfft real samples (3840 for FT8 at 12 kHz) is transformed the way kiss_fftr does it -- as an nfft/2-point complex
// FFT of the even/odd-packed signal followed by the split step with "super twiddles" -- and the complex FFT follows kiss_fft's own
// decomposition (radix 4 first, then 2, 3, 5; 1920 = 4*4*4*2*3*5) with its butterflies' exact operation order, so spectra and
// waterfall bytes are bit-identical to the CPU path.  The stage list, the input permutation and all twiddles are built on the host
// (same libm calls as kiss_fft).
//
// Execution shape (round 2): PERSISTENT CTAs walking contiguous ranges of (recording, frame) items.  Once per CTA lifetime the
// tables go to shared memory: the twiddles RE-LAID-OUT PER STAGE (tw[q k fstride] for k < m contiguous: the strided reads of the
// plain table were 16- to 32-way bank conflicts from shared memory and one sector per lane from global), the super twiddles and
// the quantiser's threshold pairs.  Per frame the windowed samples are written straight to their permuted place (inverse
// permutation table), the data is one float2 array (64-bit accesses), the butterfly index split b -> (group, k) is a multiply-high
// by a per-stage constant, and the quantiser is the straight-line form of waterfall.cu (ft8b200_selfcheck_quantiser).  The
// round-1 kernel (one CTA per frame, tables and twiddles from global memory, separate re/im arrays, runtime division per butterfly)
// took 9.4 us per 15 s recording; see profiles/ for the captures of both.
#include "common.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

using namespace ft8b200;

namespace ft8b200 {
ft8b200_ctx_t *default_ctx();
}

namespace {

constexpr int kMaxStages = 12;
constexpr int kMonThreads = 256;

struct FftPlan {
    int n;            // complex FFT length (nfft / 2)
    int nstages;
    int radix[kMaxStages], m[kMaxStages], fstride[kMaxStages];  // in EXECUTION order (innermost recursion first)
    unsigned int magic[kMaxStages];   // ceil(2^32 / m): b / m == __umulhi(b, magic) for b < 65536 (0: m == 1)
    int tw_off[kMaxStages];           // float2 offset of the stage's compact twiddles: (q - 1) * m + k -> tw[q k fstride], q = 1..radix-1
    float2 c1[kMaxStages], c2[kMaxStages];  // radix 3: c1 = tw[fstride m]; radix 5: c1 = tw[fstride m], c2 = tw[2 fstride m]
    int tw_total;
};

struct cpx { float r, i; };
__device__ __forceinline__ cpx cmul(cpx a, float2 b) {
    cpx m;
    m.r = __fsub_rn(__fmul_rn(a.r, b.x), __fmul_rn(a.i, b.y));
    m.i = __fadd_rn(__fmul_rn(a.r, b.y), __fmul_rn(a.i, b.x));
    return m;
}
__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return cpx{__fadd_rn(a.r, b.r), __fadd_rn(a.i, b.i)}; }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return cpx{__fsub_rn(a.r, b.r), __fsub_rn(a.i, b.i)}; }
// x * .5 evaluated in double then rounded to float (HALF_OF, _kiss_fft_guts.h): an exact halving
__device__ __forceinline__ float half_of(float x) { return __fmul_rn(x, 0.5f); }
__device__ __forceinline__ cpx ld(const float2 *z, int i) { const float2 v = z[i]; return cpx{v.x, v.y}; }
__device__ __forceinline__ void st(float2 *z, int i, cpx v) { z[i] = make_float2(v.r, v.i); }

// straight-line quantiser: estimate, the two thresholds around it in one 64-bit load, two compares (exactness: waterfall.cu)
__device__ __forceinline__ int quantise(float x, const float2 *__restrict__ thr2) {
    int k = (int)__fmaf_rn(6.0206f, __log2f(x), 240.0f);
    k = k < 0 ? 0 : (k > 255 ? 255 : k);
    const float2 lh = thr2[k];
    return k + (x >= lh.y ? 1 : 0) - (x < lh.x ? 1 : 0);
}

// one butterfly of radix P on z[i0 + q m], q < P, with the stage's compact twiddles tq[(q - 1) m + k]
// ix[q] = where z[i0 + q m] lives (the logical index itself, or its padded place: StaticPlan::pad)
template <int P>
__device__ __forceinline__ void butterfly(float2 *z, const float2 *tq, const int (&ix)[P], int m, int k, float2 c1, float2 c2) {
    if constexpr (P == 4) {  // kf_bfly4, kiss_fft.c:38-84
        cpx f0 = ld(z, ix[0]), f1 = ld(z, ix[1]), f2 = ld(z, ix[2]), f3 = ld(z, ix[3]);
        const cpx a = cmul(f1, tq[k]);
        const cpx bb = cmul(f2, tq[m + k]);
        const cpx c = cmul(f3, tq[2 * m + k]);
        const cpx d5 = csub(f0, bb);
        f0 = cadd(f0, bb);
        const cpx s3 = cadd(a, c), s4 = csub(a, c);
        f2 = csub(f0, s3);
        f0 = cadd(f0, s3);
        st(z, ix[0], f0);
        st(z, ix[2], f2);
        st(z, ix[1], cpx{__fadd_rn(d5.r, s4.i), __fsub_rn(d5.i, s4.r)});
        st(z, ix[3], cpx{__fsub_rn(d5.r, s4.i), __fadd_rn(d5.i, s4.r)});
    } else if constexpr (P == 2) {  // kf_bfly2, kiss_fft.c:15-36
        cpx f0 = ld(z, ix[0]), f1 = ld(z, ix[1]);
        const cpx tt = cmul(f1, tq[k]);
        f1 = csub(f0, tt);
        f0 = cadd(f0, tt);
        st(z, ix[0], f0);
        st(z, ix[1], f1);
    } else if constexpr (P == 3) {  // kf_bfly3, kiss_fft.c:86-128
        const float2 e3 = c1;
        cpx f0 = ld(z, ix[0]), f1 = ld(z, ix[1]), f2 = ld(z, ix[2]);
        const cpx s1 = cmul(f1, tq[k]);
        const cpx s2 = cmul(f2, tq[m + k]);
        const cpx s3 = cadd(s1, s2);
        cpx s0 = csub(s1, s2);
        f1.r = __fsub_rn(f0.r, half_of(s3.r));
        f1.i = __fsub_rn(f0.i, half_of(s3.i));
        s0.r = __fmul_rn(s0.r, e3.y);
        s0.i = __fmul_rn(s0.i, e3.y);
        f0 = cadd(f0, s3);
        f2.r = __fadd_rn(f1.r, s0.i);
        f2.i = __fsub_rn(f1.i, s0.r);
        f1.r = __fsub_rn(f1.r, s0.i);
        f1.i = __fadd_rn(f1.i, s0.r);
        st(z, ix[0], f0);
        st(z, ix[1], f1);
        st(z, ix[2], f2);
    } else {  // P == 5: kf_bfly5, kiss_fft.c:130-190
        const float2 ya = c1, yb = c2;
        const cpx s0 = ld(z, ix[0]);
        cpx f1 = ld(z, ix[1]), f2 = ld(z, ix[2]), f3 = ld(z, ix[3]), f4 = ld(z, ix[4]);
        const cpx s1 = cmul(f1, tq[k]);
        const cpx s2 = cmul(f2, tq[m + k]);
        const cpx s3 = cmul(f3, tq[2 * m + k]);
        const cpx s4 = cmul(f4, tq[3 * m + k]);
        const cpx s7 = cadd(s1, s4), s10 = csub(s1, s4), s8 = cadd(s2, s3), s9 = csub(s2, s3);
        st(z, ix[0], cpx{__fadd_rn(s0.r, __fadd_rn(s7.r, s8.r)), __fadd_rn(s0.i, __fadd_rn(s7.i, s8.i))});
        cpx s5, s6, s11, s12;
        s5.r = __fadd_rn(__fadd_rn(s0.r, __fmul_rn(s7.r, ya.x)), __fmul_rn(s8.r, yb.x));
        s5.i = __fadd_rn(__fadd_rn(s0.i, __fmul_rn(s7.i, ya.x)), __fmul_rn(s8.i, yb.x));
        s6.r = __fadd_rn(__fmul_rn(s10.i, ya.y), __fmul_rn(s9.i, yb.y));
        s6.i = __fsub_rn(-__fmul_rn(s10.r, ya.y), __fmul_rn(s9.r, yb.y));
        f1 = csub(s5, s6);
        f4 = cadd(s5, s6);
        s11.r = __fadd_rn(__fadd_rn(s0.r, __fmul_rn(s7.r, yb.x)), __fmul_rn(s8.r, ya.x));
        s11.i = __fadd_rn(__fadd_rn(s0.i, __fmul_rn(s7.i, yb.x)), __fmul_rn(s8.i, ya.x));
        s12.r = __fadd_rn(-__fmul_rn(s10.i, yb.y), __fmul_rn(s9.i, ya.y));
        s12.i = __fsub_rn(__fmul_rn(s10.r, yb.y), __fmul_rn(s9.r, ya.y));
        f2 = cadd(s11, s12);
        f3 = csub(s11, s12);
        st(z, ix[1], f1);
        st(z, ix[2], f2);
        st(z, ix[3], f3);
        st(z, ix[4], f4);
    }
}
// kf_bfly_generic, kiss_fft.c:192-229: any other radix p (7, 11, 13 ...: frame sizes of 11.025 / 22.05 / 44.1 kHz audio).  zg = the
// group's first element, u < m; tw = the FULL twiddle table tw[0 .. norig) in global memory (the compact per-stage tables only hold
// tw[q k fstride] for k < m).  The p inputs are read before anything is written, as the reference's scratch[] does.
constexpr int kMaxRadix = 32;
__device__ __noinline__ void butterfly_generic(float2 *zg, const float2 *__restrict__ tw, int u, int m, int p, int fstride, int norig) {
    cpx scratch[kMaxRadix];
    int k = u;
    for (int q1 = 0; q1 < p; ++q1) { scratch[q1] = ld(zg, k); k += m; }
    k = u;
    for (int q1 = 0; q1 < p; ++q1) {
        int twidx = 0;
        cpx acc = scratch[0];
        for (int q = 1; q < p; ++q) {
            twidx += fstride * k;
            if (twidx >= norig) twidx -= norig;
            const cpx tt = cmul(scratch[q], __ldg(tw + twidx));
            acc.r = __fadd_rn(acc.r, tt.r);   // C_ADDTO
            acc.i = __fadd_rn(acc.i, tt.i);
        }
        st(zg, k, acc);
        k += m;
    }
}

template <int P>
__device__ __forceinline__ void butterfly_at(float2 *z, const float2 *tq, int i0, int m, int k, float2 c1, float2 c2) {
    int ix[P];
#pragma unroll
    for (int q = 0; q < P; ++q) ix[q] = i0 + q * m;
    butterfly<P>(z, tq, ix, m, k, c1, c2);
}

// Compile-time stage lists of the two geometries in use (kf_factor's decomposition, execution order = innermost recursion first):
// FT8 at 12 kHz: 3840-point frames -> 1920 = 5 * 3 * 2 * 4 * 4 * 4;  FT4 at 12 kHz: 1152-point frames -> 576 = 3 * 3 * 4 * 4 * 4.
// The launcher uses them only when the host-built plan (build_plan) is identical; any other size runs the generic stage loop.
// Shared-memory placement (measured in round 2: 47 M bank conflicts per 128 recordings, twice the ideal wavefront count, on a
// kernel whose shared-memory pipe was ~70 % busy).  Two compile-time knobs per plan, chosen by exhaustive simulation of the
// 16 x 8-byte banks (tools/fft_bank_sim.py): `pad` -- one unused float2 after every `pad` elements (element i lives at i + i / pad) --
// and, per stage, which index runs fastest across a warp's butterflies: k (the twiddle index; consecutive elements) or g (the group;
// stride radix * m, odd after padding).  3840-point frames: pad 240 + {k, g, g, g, k, k} -> 1702 wavefronts per frame against an
// ideal 1682 (was 3370, the scattered leaf store alone 960).
template <int N> struct StaticPlan { static constexpr int ns = 0; static constexpr int pad = 0; };
static_assert(StaticPlan<0>::ns == 0 && StaticPlan<0>::pad == 0, "the generic kernel has no compile-time plan");
template <> struct StaticPlan<1920> {
    static constexpr int ns = 6;
    static constexpr int pad = 240;
    __host__ __device__ static constexpr int radix(int s) { constexpr int r[6] = {5, 3, 2, 4, 4, 4}; return r[s]; }
    __host__ __device__ static constexpr int m(int s) { constexpr int v[6] = {1, 5, 15, 30, 120, 480}; return v[s]; }
    __host__ __device__ static constexpr bool gfast(int s) { constexpr bool v[6] = {false, true, true, true, false, false}; return v[s]; }
};
template <> struct StaticPlan<576> {
    static constexpr int ns = 5;
    static constexpr int pad = 0;
    __host__ __device__ static constexpr int radix(int s) { constexpr int r[5] = {3, 3, 4, 4, 4}; return r[s]; }
    __host__ __device__ static constexpr int m(int s) { constexpr int v[5] = {1, 3, 9, 36, 144}; return v[s]; }
    __host__ __device__ static constexpr bool gfast(int s) { constexpr bool v[5] = {false, true, false, false, false}; return v[s]; }
};
template <int N> __host__ __device__ constexpr int padded(int i) { return StaticPlan<N>::pad > 0 ? i + i / StaticPlan<N>::pad : i; }
// elements of z: n plus the pad slots
// (rounded up to an even count: the arrays behind z keep their 16-byte alignment)
__host__ __device__ constexpr int z_len(int static_n, int n) { return static_n == 1920 ? (padded<1920>(1920 - 1) + 2) & ~1 : n; }
template <int N> __host__ __device__ constexpr int static_tw_off(int s) { return s == 0 ? 0 : static_tw_off<N>(s - 1) + (StaticPlan<N>::radix(s - 1) - 1) * StaticPlan<N>::m(s - 1); }

template <int N, int S>
__device__ __forceinline__ void static_stages(float2 *z, const float2 *tws, const FftPlan &plan, int t) {
    if constexpr (S < StaticPlan<N>::ns) {
        constexpr int p = StaticPlan<N>::radix(S), m = StaticPlan<N>::m(S), nbf = N / p, groups = nbf / m, B = StaticPlan<N>::pad;
        const float2 *tq = tws + static_tw_off<N>(S);
#pragma unroll
        for (int b0 = 0; b0 < nbf; b0 += kMonThreads) {
            const int b = b0 + t;
            if (b0 + kMonThreads <= nbf || b < nbf) {
                // m and groups are compile-time constants: multiply + shift
                const int g = StaticPlan<N>::gfast(S) ? b % groups : b / m, k = StaticPlan<N>::gfast(S) ? b / groups : b - g * m;
                const int i0 = g * p * m + k;
                int ix[p];
                if constexpr (B == 0) {
#pragma unroll
                    for (int q = 0; q < p; ++q) ix[q] = i0 + q * m;
                } else if constexpr (B % (p * m) == 0) {          // the butterfly's p legs share a pad block
                    const int off = g / (B / (p * m));
#pragma unroll
                    for (int q = 0; q < p; ++q) ix[q] = i0 + q * m + off;
                } else if constexpr ((p * m) % B == 0 && B % m == 0) {   // B / m legs per pad block (k < m <= B)
                    const int off = g * (p * m / B);
#pragma unroll
                    for (int q = 0; q < p; ++q) ix[q] = i0 + q * m + off + q / (B / m);
                } else {                                           // m a multiple of B: each leg spans m / B blocks
                    static_assert(m % B == 0 && (p * m) % B == 0, "pad block must nest with the stage geometry");
                    const int off = g * (p * m / B) + k / B;
#pragma unroll
                    for (int q = 0; q < p; ++q) ix[q] = i0 + q * m + off + q * (m / B);
                }
                butterfly<p>(z, tq, ix, m, k, plan.c1[S], plan.c2[S]);
            }
        }
        __syncthreads();
        static_stages<N, S + 1>(z, tws, plan, t);
    }
}

// dynamic smem: z float2[z_len] | stage twiddles float2[tw_total] | super twiddles float2[n/2 + 1] | thr2 float2[256] | out u8[2 n]
// kN = 0: generic stage loop from the run-time plan; kN = 1920 / 576: the stage loop unrolled at compile time (StaticPlan)
template <int kN>
__global__ void __launch_bounds__(kMonThreads, 4)
monitor_frames_kernel(const float *__restrict__ audio, size_t slot_stride, int n_samples, long first_start, int hop, int nfft, FftPlan plan,
                      const uint16_t *__restrict__ inv_perm, const float2 *__restrict__ tw_stage, const float2 *__restrict__ super_tw,
                      const float *__restrict__ wnorm, const float *__restrict__ thr_g, int num_bins, int freq_osr, int n_frames, int total_items,
                      uint8_t *__restrict__ mag, size_t mag_slot_stride, unsigned int *__restrict__ xmax_bits) {
    extern __shared__ __align__(16) float2 smem2[];
    const int n = kN > 0 ? kN : plan.n, t = threadIdx.x;
    float2 *z = smem2, *tws = z + z_len(kN, n), *sup = tws + plan.tw_total, *thr2 = sup + (n / 2 + 1);
    uint8_t *outb = reinterpret_cast<uint8_t *>(thr2 + 256);
    const int it_begin = (int)((long long)blockIdx.x * total_items / gridDim.x), it_end = (int)((long long)(blockIdx.x + 1) * total_items / gridDim.x);
    if (it_begin >= it_end) return;
    for (int k = t; k < plan.tw_total; k += kMonThreads) tws[k] = tw_stage[k];
    for (int k = t; k < n / 2 + 1; k += kMonThreads) sup[k] = super_tw[k];
    for (int k = t; k < 256; k += kMonThreads) thr2[k] = make_float2(thr_g[k], thr_g[k + 1]);
    const int wanted = num_bins * freq_osr;  // bins 0 .. wanted-1 are stored (wanted <= n)
    const bool osr2 = freq_osr == 2;         // the usual oversampling: no run-time division in the store index
    auto cell = [&](int k) { return osr2 ? (k & 1) * num_bins + (k >> 1) : (k % freq_osr) * num_bins + k / freq_osr; };

    // Static plans: a thread's sample pairs of the NEXT frame are fetched into registers as soon as the current frame's have been
    // windowed into z, so the loads are in flight during the whole transform (waiting for them at the top of every frame was 11 % of
    // the kernel's stall samples).
    constexpr int kPer = kN > 0 ? (kN + kMonThreads - 1) / kMonThreads : 1;
    float2 pre[kPer];
    auto fetch = [&](int item) {
        const int slot = item / n_frames, frame = item - slot * n_frames;
        const float *x = audio + (size_t)slot * slot_stride;
        const long start = first_start + (long)frame * hop;
        const bool pairs = ((((size_t)x) >> 2) + (size_t)(start & 1)) % 2 == 0;   // x + start is 8-byte aligned (x is a float pointer)
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            const int src = t + j * kMonThreads;
            const long p0 = start + 2 * src, p1 = p0 + 1;
            float2 v = make_float2(0.0f, 0.0f);
            if (kPer * kMonThreads == n || src < n) {
                if (pairs && p0 >= 0 && p1 < n_samples) {
                    v = *reinterpret_cast<const float2 *>(x + p0);
                } else {
                    if (p0 >= 0 && p0 < n_samples) v.x = x[p0];
                    if (p1 >= 0 && p1 < n_samples) v.y = x[p1];
                }
            }
            pre[j] = v;
        }
    };
    if constexpr (kN > 0) fetch(it_begin);

    for (int item = it_begin; item < it_end; ++item) {
        const int slot = item / n_frames, frame = item - slot * n_frames;
        __syncthreads();  // tables loaded / the previous frame's bytes have left outb and z
        // windowed, normalised frame, packed (even, odd) -> complex, written to kiss_fft's permuted leaf place
        if constexpr (kN > 0) {
#pragma unroll
            for (int j = 0; j < kPer; ++j) {
                const int src = t + j * kMonThreads;
                if (kPer * kMonThreads == n || src < n) {
                    const float2 w = *reinterpret_cast<const float2 *>(wnorm + 2 * src);
                    z[inv_perm[src]] = make_float2(__fmul_rn(w.x, pre[j].x), __fmul_rn(w.y, pre[j].y));   // (fft_norm * window[pos]) * last_frame[pos], decode_ft8.c:191; inv_perm holds the (padded) place
                }
            }
            if (item + 1 < it_end) fetch(item + 1);
        } else {
            const float *x = audio + (size_t)slot * slot_stride;
            const long start = first_start + (long)frame * hop;
            for (int src = t; src < n; src += kMonThreads) {
                const long p0 = start + 2 * src, p1 = p0 + 1;
                const float a = (p0 >= 0 && p0 < n_samples) ? x[p0] : 0.0f;
                const float b = (p1 >= 0 && p1 < n_samples) ? x[p1] : 0.0f;
                const float2 w = *reinterpret_cast<const float2 *>(wnorm + 2 * src);
                z[inv_perm[src]] = make_float2(__fmul_rn(w.x, a), __fmul_rn(w.y, b));
            }
        }
        __syncthreads();
        if constexpr (kN > 0) {
            static_stages<kN, 0>(z, tws, plan, t);
        } else {
            for (int s = 0; s < plan.nstages; ++s) {
                const int p = plan.radix[s], m = plan.m[s];
                const unsigned int magic = plan.magic[s];
                const float2 *tq = tws + plan.tw_off[s];
                const int nbf = n / p;  // butterflies in this stage
                for (int b = t; b < nbf; b += kMonThreads) {
                    const int g = magic ? (int)__umulhi((unsigned int)b, magic) : b, k = b - g * m;
                    const int i0 = g * p * m + k;
                    if (p == 4) butterfly_at<4>(z, tq, i0, m, k, plan.c1[s], plan.c2[s]);
                    else if (p == 2) butterfly_at<2>(z, tq, i0, m, k, plan.c1[s], plan.c2[s]);
                    else if (p == 3) butterfly_at<3>(z, tq, i0, m, k, plan.c1[s], plan.c2[s]);
                    else if (p == 5) butterfly_at<5>(z, tq, i0, m, k, plan.c1[s], plan.c2[s]);
                    else butterfly_generic(z + g * p * m, tw_stage + plan.tw_total, k, m, p, plan.fstride[s], n);   // the full table follows the compact ones
                }
                __syncthreads();
            }
        }
        // real-FFT split (kiss_fftr.c:80-113) fused with |X|^2 -> dB -> uint8 (decode_ft8.c:196-213)
        float xmax = 0.0f;
        for (int k = t; k <= n / 2; k += kMonThreads) {
            cpx lo, hi;  // bins k and n-k
            bool have_hi = false;
            if (k == 0) {
                const float2 z0 = z[0];   // element 0 is never displaced by the padding
                lo.r = __fadd_rn(z0.x, z0.y);
                lo.i = 0.0f;
            } else {
                const cpx fpk = ld(z, padded<kN>(k));
                const float2 zn = z[padded<kN>(n - k)];
                const cpx fpnk{zn.x, -zn.y};
                const cpx f1 = cadd(fpk, fpnk), f2 = csub(fpk, fpnk);
                const cpx tt = cmul(f2, sup[k - 1]);
                lo.r = half_of(__fadd_rn(f1.r, tt.r));
                lo.i = half_of(__fadd_rn(f1.i, tt.i));
                hi.r = half_of(__fsub_rn(f1.r, tt.r));
                hi.i = half_of(__fsub_rn(tt.i, f1.i));
                have_hi = true;
                if (k == n - k) lo = hi;  // the loop's last iteration writes freqdata[k] then freqdata[ncfft-k]: the latter wins
            }
            if (k < wanted) {
                const float m2 = __fadd_rn(__fmul_rn(lo.i, lo.i), __fmul_rn(lo.r, lo.r));
                const float xv = __fadd_rn(1E-12f, m2);
                xmax = fmaxf(xmax, xv);
                outb[cell(k)] = (uint8_t)quantise(xv, thr2);
            }
            const int kh = n - k;
            if (have_hi && kh != k && kh < wanted) {
                const float m2 = __fadd_rn(__fmul_rn(hi.i, hi.i), __fmul_rn(hi.r, hi.r));
                const float xv = __fadd_rn(1E-12f, m2);
                xmax = fmaxf(xmax, xv);
                outb[cell(kh)] = (uint8_t)quantise(xv, thr2);
            }
        }
        if (xmax_bits) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
            if ((t & 31) == 0) atomicMax(xmax_bits + slot, __float_as_uint(xmax));
        }
        __syncthreads();
        uint8_t *out = mag + (size_t)slot * mag_slot_stride + (size_t)frame * wanted;
        if (((((size_t)out) | (size_t)wanted) & 15) == 0) {
            for (int k = t; k < wanted / 16; k += kMonThreads) reinterpret_cast<uint4 *>(out)[k] = reinterpret_cast<const uint4 *>(outb)[k];
        } else {
            for (int k = t; k < wanted; k += kMonThreads) out[k] = outb[k];
        }
    }
}

// ---- host-side plan / tables (same libm calls as kiss_fft / kiss_fftr / decode_ft8.c) ----------
struct MonTables {
    int nfft = 0;
    int device = 0;
    FftPlan plan = {};
    uint16_t *d_perm = nullptr;
    float2 *d_tw = nullptr, *d_super = nullptr;
    float *d_wnorm = nullptr;
    std::vector<float> window;  // Hann, host copy
    float fft_norm = 0;
    int static_n = 0;           // 1920 / 576: the plan is the one StaticPlan spells out (compile-time stage list, padded placement); else 0
    size_t smem = 0;            // launch shape of monitor_frames_kernel for these tables, worked out once
    long max_grid = 0;
};

bool build_plan(int n, FftPlan &plan, std::vector<uint16_t> &perm) {
    // kf_factor (kiss_fft.c:303-324): radices in RECURSION order, outermost first
    int radix[kMaxStages], rem[kMaxStages], nf = 0;
    int left = n, q = 4;
    const double lim = floor(sqrt((double)n));
    do {
        while (left % q) {
            q = (q == 4) ? 2 : (q == 2) ? 3 : q + 2;
            if (q > lim) q = left;
        }
        left /= q;
        if (nf >= kMaxStages || q > kMaxRadix) return false;   // 2, 3, 4, 5 have their own butterflies, anything else up to kMaxRadix the generic one
        radix[nf] = q;
        rem[nf] = left;
        ++nf;
    } while (left > 1);
    // execution order = innermost first; fstride of level L = product of the radices above it
    plan.n = n;
    plan.nstages = nf;
    int stride_above[kMaxStages];
    int acc = 1;
    for (int l = 0; l < nf; ++l) { stride_above[l] = acc; acc *= radix[l]; }
    for (int s = 0; s < nf; ++s) {
        const int l = nf - 1 - s;
        plan.radix[s] = radix[l];
        plan.m[s] = rem[l];
        plan.fstride[s] = stride_above[l];
    }
    // leaf copy order (kf_work, kiss_fft.c:273-289): output index o <- input index sum(digit_l * stride_above[l])
    perm.resize((size_t)n);
    for (int o = 0; o < n; ++o) {
        int r = o, src = 0;
        for (int l = 0; l < nf; ++l) {
            const int d = r / rem[l];
            r -= d * rem[l];
            src += d * stride_above[l];
        }
        perm[(size_t)o] = (uint16_t)src;
    }
    return true;
}

std::mutex g_mon_mu;
std::vector<MonTables *> g_tables;

// tables are device allocations: one set per (device, nfft); the caller has made `device` current
MonTables *get_tables(int device, int nfft) {
    std::lock_guard<std::mutex> lk(g_mon_mu);
    for (MonTables *t : g_tables) if (t->nfft == nfft && t->device == device) return t;
    if (nfft < 8 || (nfft & 1) || nfft / 2 > 65535) return nullptr;
    MonTables *t = new MonTables();
    t->nfft = nfft;
    t->device = device;
    const int n = nfft / 2;
    std::vector<uint16_t> perm;
    if (!build_plan(n, t->plan, perm)) { delete t; return nullptr; }
    std::vector<float2> tw((size_t)n), sup((size_t)(n / 2 + 1));
    build_twiddles(n, tw.data());
    // per-stage compact twiddles tw[q k fstride] (q = 1..radix-1, k < m), the radix-3/5 constants, the division constants,
    // and the INVERSE of the leaf permutation (the kernel scatters input sample pair `src` to its permuted place)
    std::vector<float2> tws;
    for (int s2 = 0; s2 < t->plan.nstages; ++s2) {
        const int p = t->plan.radix[s2], m = t->plan.m[s2], fs = t->plan.fstride[s2];
        t->plan.tw_off[s2] = (int)tws.size();
        if (p <= 5)   // the generic butterfly reads the full table instead
            for (int q = 1; q < p; ++q)
                for (int k = 0; k < m; ++k) tws.push_back(tw[(size_t)(q * k * fs)]);
        t->plan.magic[s2] = m > 1 ? (unsigned int)((0x100000000ull + (unsigned long long)m - 1) / (unsigned long long)m) : 0u;
        t->plan.c1[s2] = tw[(size_t)(fs * m) % (size_t)n];
        t->plan.c2[s2] = tw[(size_t)(2 * fs * m) % (size_t)n];
    }
    t->plan.tw_total = (int)tws.size();
    bool any_generic = false;
    for (int s2 = 0; s2 < t->plan.nstages; ++s2) any_generic |= t->plan.radix[s2] > 5;
    if (any_generic) tws.insert(tws.end(), tw.begin(), tw.end());   // behind the compact tables: what butterfly_generic indexes (global memory only)
    // compile-time stage list when the host-built plan is exactly the one StaticPlan spells out
    auto matches = [&](auto tag) {
        using SP = decltype(tag);
        if (SP::ns == 0 || t->plan.nstages != SP::ns) return false;
        for (int s2 = 0; s2 < SP::ns; ++s2) if (t->plan.radix[s2] != SP::radix(s2) || t->plan.m[s2] != SP::m(s2)) return false;
        return true;
    };
    t->static_n = (n == 1920 && matches(StaticPlan<1920>{})) ? 1920 : (n == 576 && matches(StaticPlan<576>{})) ? 576 : 0;
    std::vector<uint16_t> inv((size_t)n);
    for (int o = 0; o < n; ++o) inv[perm[(size_t)o]] = (uint16_t)(t->static_n == 1920 ? padded<1920>(o) : o);   // the element's place in z
    perm.swap(inv);
    for (int k = 0; k < n / 2; ++k) {  // kiss_fftr_alloc, kiss_fftr.c:50-56
        const double phase = -3.14159265358979323846264338327 * ((double)(k + 1) / n + .5);
        sup[(size_t)k].x = (float)cos(phase);
        sup[(size_t)k].y = (float)sin(phase);
    }
    t->fft_norm = 2.0f / nfft;  // decode_ft8.c:120
    t->window.resize((size_t)nfft);
    std::vector<float> wn((size_t)nfft);
    for (int i = 0; i < nfft; ++i) {  // hann_i, decode_ft8.c:35-39
        const float x = sinf((float)M_PI * i / nfft);
        t->window[(size_t)i] = x * x;
        wn[(size_t)i] = t->fft_norm * t->window[(size_t)i];
    }
    bool ok = cudaMalloc(&t->d_perm, sizeof(uint16_t) * n) == cudaSuccess && cudaMalloc(&t->d_tw, sizeof(float2) * (tws.size() + 1)) == cudaSuccess &&
              cudaMalloc(&t->d_super, sizeof(float2) * (n / 2 + 1)) == cudaSuccess && cudaMalloc(&t->d_wnorm, sizeof(float) * nfft) == cudaSuccess &&
              cudaMemcpy(t->d_perm, perm.data(), sizeof(uint16_t) * n, cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(t->d_tw, tws.data(), sizeof(float2) * tws.size(), cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(t->d_super, sup.data(), sizeof(float2) * (n / 2 + 1), cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(t->d_wnorm, wn.data(), sizeof(float) * nfft, cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) { delete t; return nullptr; }
    g_tables.push_back(t);
    return t;
}

constexpr int kMaxDevices = 64;
float *g_thr[kMaxDevices] = {};  // per device: copy of the dB step thresholds (shared by all monitors on that device)
float *thresholds(int device) {
    if (device < 0 || device >= kMaxDevices) return nullptr;
    std::lock_guard<std::mutex> lk(g_mon_mu);
    if (!g_thr[device]) {
        float t[257];
        build_db_thresholds(t);
        if (cudaMalloc(&g_thr[device], sizeof(t)) != cudaSuccess || cudaMemcpy(g_thr[device], t, sizeof(t), cudaMemcpyHostToDevice) != cudaSuccess)
            g_thr[device] = nullptr;
    }
    return g_thr[device];
}

cudaError_t launch_frames(const MonTables *t, const float *d_audio, size_t slot_stride, int n_samples, long first_start, int hop, int n_frames,
                          int n_slots, int num_bins, int freq_osr, uint8_t *d_mag, size_t mag_slot_stride, unsigned int *d_xmax, cudaStream_t st) {
    const int n = t->plan.n;
    cudaError_t e;
    auto kern = monitor_frames_kernel<0>;
    if (t->static_n == 1920) kern = monitor_frames_kernel<1920>;
    else if (t->static_n == 576) kern = monitor_frames_kernel<576>;
    if (t->max_grid == 0) {  // once per (device, nfft): monitor_process() launches this 93 times per recording
        // z | stage twiddles | super twiddles | threshold pairs | one frame's bytes (see monitor_frames_kernel)
        const size_t need = sizeof(float2) * ((size_t)z_len(t->static_n, n) + (size_t)t->plan.tw_total + (size_t)(n / 2 + 1) + 256) + (((size_t)2 * n + 15) & ~(size_t)15);
        int sms = 0, per_sm = 0, optin = 0;
        if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, t->device)) != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, t->device)) != cudaSuccess) return e;
        if (need > (size_t)optin) return cudaErrorInvalidConfiguration;   // the frame does not fit one CTA's shared memory: refused, no CUDA call fails
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need)) != cudaSuccess) return e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kMonThreads, need)) != cudaSuccess) return e;
        MonTables *mt = const_cast<MonTables *>(t);
        mt->smem = need;
        mt->max_grid = (long)sms * (per_sm > 0 ? per_sm : 1);
    }
    const size_t smem = t->smem;
    // the opt-in is per FUNCTION, not per table set: another geometry may have lowered it since
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    const long total = (long)n_frames * n_slots;
    long grid = t->max_grid;   // persistent: every CTA walks a contiguous range of (recording, frame) items
    if (grid > total) grid = total;
    if (total >= (1l << 31)) return cudaErrorInvalidValue;
    kern<<<(unsigned)grid, kMonThreads, smem, st>>>(d_audio, slot_stride, n_samples, first_start, hop, t->nfft, t->plan, t->d_perm, t->d_tw,
                                                                      t->d_super, t->d_wnorm, thresholds(t->device), num_bins, freq_osr, n_frames, (int)total,
                                                                      d_mag, mag_slot_stride, d_xmax);
    return cudaGetLastError();
}

struct MonGeom { int block_size, subblock_size, nfft, max_blocks, num_bins; float symbol_period; };
MonGeom geometry(int sample_rate, int time_osr, int freq_osr, int protocol) {
    MonGeom g;
    const float slot_time = (protocol == PROTO_FT4) ? 7.5f : 15.0f;       // constants.h:12-16
    g.symbol_period = (protocol == PROTO_FT4) ? 0.048f : 0.160f;
    g.block_size = (int)(sample_rate * g.symbol_period);                   // decode_ft8.c:116-119
    g.subblock_size = g.block_size / time_osr;
    g.nfft = g.block_size * freq_osr;
    g.max_blocks = (int)(slot_time / g.symbol_period);                     // decode_ft8.c:144-145
    g.num_bins = (int)(sample_rate * g.symbol_period / 2);
    return g;
}

struct MonitorDev {  // hangs off monitor_t.fft_work
    const MonTables *tables;
    float *d_buf;          // history (nfft - subblock floats) + up to max_blocks new blocks
    uint8_t *d_mag;        // up to max_blocks blocks' worth of bytes
    unsigned int *d_xmax;
    float *h_buf;          // pinned, same shape as d_buf
    cudaStream_t st;
    // deferred mode (ft8b200_monitor_set_deferred): monitor_process() only appends the block to h_buf; the frames are transformed in
    // ONE launch when the waterfall is needed -- ft8_find_sync / ft8_decode on this monitor's wf, ft8b200_monitor_flush,
    // monitor_reset, monitor_free -- instead of a copy-launch-copy-synchronise round trip per 160 ms block
    bool deferred;
    int pending;           // blocks appended since the last flush
    int max_blocks;
};

// monitors in deferred mode, by the host address of their waterfall bytes (what ft8_find_sync / ft8_decode are handed)
std::mutex g_deferred_mu;
std::vector<monitor_t *> g_deferred;

void die(const char *what) {
    fprintf(stderr, "libft8b200: monitor: %s (%s)\n", what, cudaGetErrorString(cudaGetLastError()));
    abort();
}

}  // namespace

namespace ft8b200 {
// ft8_find_sync / ft8_decode (dropin.cu) call this with the waterfall they were handed: a deferred monitor that owns those
// bytes transforms its pending blocks first
void monitor_flush_for_mag(const uint8_t *mag) {
    monitor_t *hit = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_deferred_mu);
        for (monitor_t *m : g_deferred) if (m->wf.mag == mag) { hit = m; break; }
    }
    if (hit) ft8b200_monitor_flush(hit);
}
}  // namespace ft8b200

extern "C" {

int ft8b200_monitor_waterfall(ft8b200_ctx_t *ctx, const float *d_audio, size_t slot_stride_samples, int n_samples, int n_slots, int sample_rate,
                              int time_osr, int freq_osr, int protocol, uint8_t *d_mag, size_t mag_slot_stride, int *num_blocks_out, void *stream) {
    if (!ctx || !d_audio || !d_mag || n_slots < 1 || n_samples < 0 || time_osr < 1 || freq_osr < 1) return api_error(FT8B200_EINVAL, "ft8b200_monitor_waterfall: bad argument");
    const MonGeom g = geometry(sample_rate, time_osr, freq_osr, protocol);
    if (g.block_size < 2 || g.subblock_size * time_osr != g.block_size)
        return api_error(FT8B200_EINVAL, "ft8b200_monitor_waterfall: the block of one symbol period does not divide by time_osr at this sample rate");
    const int device = ctx_device(ctx);
    if (cudaSetDevice(device) != cudaSuccess) return FT8B200_CUDA_FAIL();  // tables and launches belong to the context's device, not the caller's current one
    const MonTables *t = get_tables(device, g.nfft);
    if (!t || !thresholds(device)) {
        char why[160];
        snprintf(why, sizeof(why), "ft8b200_monitor_waterfall: unsupported frame size %d (odd, a prime factor above %d, or more than 131070 points)", g.nfft, kMaxRadix);
        return api_error(FT8B200_EINVAL, why);
    }
    int nb = n_samples / g.block_size;
    if (nb > g.max_blocks) nb = g.max_blocks;
    if (num_blocks_out) *num_blocks_out = nb;
    if (nb == 0) return 0;
    const size_t stride = (size_t)time_osr * freq_osr * g.num_bins;
    if (mag_slot_stride < (size_t)nb * stride) return api_error(FT8B200_EINVAL, "ft8b200_monitor_waterfall: mag_slot_stride is smaller than one recording's waterfall");
    cudaStream_t st = stream ? (cudaStream_t)stream : (cudaStream_t)ft8b200_cuda_stream(ctx);
    // frame f ends at sample (f+1)*subblock; the reference's last_frame starts out as (zeroed) history
    cudaError_t e = launch_frames(t, d_audio, slot_stride_samples, n_samples, (long)g.subblock_size - g.nfft, g.subblock_size, nb * time_osr, n_slots,
                                  g.num_bins, freq_osr, d_mag, mag_slot_stride, nullptr, st);
    if (e == cudaErrorInvalidConfiguration) {
        char why[160];
        snprintf(why, sizeof(why), "ft8b200_monitor_waterfall: a frame of %d points (sample rate %d) does not fit one CTA's shared memory", g.nfft, sample_rate);
        return ::ft8b200::cuda_error(e, why);
    }
    return e == cudaSuccess ? 0 : ::ft8b200::cuda_error(e, __func__);
}

void monitor_init(monitor_t *me, const monitor_config_t *cfg) {
    ft8b200_ctx_t *ctx = default_ctx();
    const MonGeom g = geometry(cfg->sample_rate, cfg->time_osr, cfg->freq_osr, (int)cfg->protocol);
    const int device = ctx_device(ctx);
    if (cudaSetDevice(device) != cudaSuccess) die("monitor_init: cudaSetDevice");
    const MonTables *t = get_tables(device, g.nfft);
    if (!t || !thresholds(device)) {
        fprintf(stderr, "libft8b200: monitor_init: unsupported FFT size %d (a prime factor above %d, or more than 65535 points)\n", g.nfft, 32);
        abort();
    }
    me->symbol_period = g.symbol_period;
    me->block_size = g.block_size;
    me->subblock_size = g.subblock_size;
    me->nfft = g.nfft;
    me->fft_norm = t->fft_norm;
    me->window = (float *)malloc(sizeof(float) * (size_t)g.nfft);
    memcpy(me->window, t->window.data(), sizeof(float) * (size_t)g.nfft);
    // the reference leaves last_frame uninitialised (decode_ft8.c:131); zero is the deterministic choice
    me->last_frame = (float *)calloc((size_t)g.nfft, sizeof(float));
    waterfall_init(&me->wf, g.max_blocks, g.num_bins, cfg->time_osr, cfg->freq_osr);
    me->wf.protocol = cfg->protocol;
    me->max_mag = -120.0f;
    MonitorDev *d = new MonitorDev();
    d->tables = t;
    d->st = (cudaStream_t)ft8b200_cuda_stream(ctx);
    d->deferred = false;
    d->pending = 0;
    d->max_blocks = g.max_blocks;
    const size_t nbuf = (size_t)g.nfft - g.subblock_size + (size_t)g.block_size * (size_t)g.max_blocks;
    if (cudaMalloc(&d->d_buf, sizeof(float) * nbuf) != cudaSuccess || cudaMalloc(&d->d_mag, (size_t)me->wf.block_stride * (size_t)g.max_blocks) != cudaSuccess ||
        cudaMalloc(&d->d_xmax, sizeof(unsigned int)) != cudaSuccess || cudaHostAlloc(&d->h_buf, sizeof(float) * nbuf, cudaHostAllocDefault) != cudaSuccess)
        die("monitor_init: allocation failed");
    me->fft_work = d;
    me->fft_cfg = nullptr;
    const char *env = getenv("FT8B200_MONITOR_DEFERRED");
    if (env && atoi(env) != 0) ft8b200_monitor_set_deferred(me, 1);
}

// transform the `blocks` blocks that sit behind the history in h_buf, write their bytes to wf.mag from block `first_block` on
static void run_blocks(monitor_t *me, MonitorDev *d, int first_block, int blocks) {
    const int nfft = me->nfft, sub = me->subblock_size, blk = me->block_size, tosr = me->wf.time_osr;
    const size_t nhist = (size_t)nfft - sub, nbuf = nhist + (size_t)blk * blocks;
    const size_t off = (size_t)first_block * me->wf.block_stride, bytes = (size_t)blocks * me->wf.block_stride;
    unsigned int xbits = 0;
    bool ok = cudaMemcpyAsync(d->d_buf, d->h_buf, sizeof(float) * nbuf, cudaMemcpyHostToDevice, d->st) == cudaSuccess &&
              cudaMemsetAsync(d->d_xmax, 0, sizeof(unsigned int), d->st) == cudaSuccess &&
              launch_frames(d->tables, d->d_buf, 0, (int)nbuf, 0, sub, tosr * blocks, 1, me->wf.num_bins, me->wf.freq_osr, d->d_mag, 0, d->d_xmax, d->st) == cudaSuccess &&
              cudaMemcpyAsync(me->wf.mag + off, d->d_mag, bytes, cudaMemcpyDeviceToHost, d->st) == cudaSuccess &&
              cudaMemcpyAsync(&xbits, d->d_xmax, sizeof(xbits), cudaMemcpyDeviceToHost, d->st) == cudaSuccess && cudaStreamSynchronize(d->st) == cudaSuccess;
    if (!ok) die("monitor_process failed");
    // host copy of the sliding frame, as the reference keeps it (decode_ft8.c:178-186): the last nfft samples seen
    memcpy(me->last_frame, d->h_buf + nbuf - (size_t)nfft, sizeof(float) * (size_t)nfft);
    float xmax;
    memcpy(&xmax, &xbits, 4);
    const float db = 10.0f * log10f(xmax);  // max over the blocks of 10*log10f(x): log10f is monotone, so this is their max dB
    if (db > me->max_mag) me->max_mag = db;
}

void monitor_free(monitor_t *me) {
    MonitorDev *d = (MonitorDev *)me->fft_work;
    if (d && d->deferred) ft8b200_monitor_set_deferred(me, 0);  // flushes and leaves the registry
    if (d) {
        cudaSetDevice(d->tables->device);
        cudaStreamSynchronize(d->st);
        cudaFree(d->d_buf); cudaFree(d->d_mag); cudaFree(d->d_xmax); cudaFreeHost(d->h_buf);
        delete d;
    }
    waterfall_free(&me->wf);
    free(me->last_frame);
    free(me->window);
    me->fft_work = nullptr;
}

void monitor_reset(monitor_t *me) {  // decode_ft8.c:220-224: last_frame is NOT cleared
    ft8b200_monitor_flush(me);       // deferred blocks still shape last_frame (and max_mag is reset below, after them)
    me->wf.num_blocks = 0;
    me->max_mag = 0;
}

void monitor_process(monitor_t *me, const float *frame) {
    if (me->wf.num_blocks >= me->wf.max_blocks) return;  // silent no-op once full (decode_ft8.c:165-166)
    MonitorDev *d = (MonitorDev *)me->fft_work;
    const int nfft = me->nfft, sub = me->subblock_size, blk = me->block_size;
    const size_t nhist = (size_t)nfft - sub;
    if (d->deferred) {  // append only; the device work happens at the flush
        if (d->pending == 0) memcpy(d->h_buf, me->last_frame + sub, sizeof(float) * nhist);
        memcpy(d->h_buf + nhist + (size_t)d->pending * blk, frame, sizeof(float) * (size_t)blk);
        ++d->pending;
        ++me->wf.num_blocks;
        return;
    }
    if (cudaSetDevice(d->tables->device) != cudaSuccess) die("monitor_process: cudaSetDevice");  // the caller's thread may have another device current
    memcpy(d->h_buf, me->last_frame + sub, sizeof(float) * nhist);
    memcpy(d->h_buf + nhist, frame, sizeof(float) * (size_t)blk);
    run_blocks(me, d, me->wf.num_blocks, 1);
    ++me->wf.num_blocks;
}

// Deferred mode: see MonitorDev.  What a caller must know: between a monitor_process() and the next flush, wf.mag and max_mag do
// not yet reflect the appended blocks (wf.num_blocks does).  ft8_find_sync() and ft8_decode() flush by themselves.
int ft8b200_monitor_flush(monitor_t *me) {
    MonitorDev *d = me ? (MonitorDev *)me->fft_work : nullptr;
    if (!d) return FT8B200_BAD_ARG();
    if (d->pending == 0) return 0;
    if (cudaSetDevice(d->tables->device) != cudaSuccess) die("ft8b200_monitor_flush: cudaSetDevice");
    const int blocks = d->pending;
    d->pending = 0;
    run_blocks(me, d, me->wf.num_blocks - blocks, blocks);
    return blocks;
}

int ft8b200_monitor_set_deferred(monitor_t *me, int on) {
    MonitorDev *d = me ? (MonitorDev *)me->fft_work : nullptr;
    if (!d) return FT8B200_BAD_ARG();
    if (!on) ft8b200_monitor_flush(me);
    std::lock_guard<std::mutex> lk(g_deferred_mu);
    for (size_t k = 0; k < g_deferred.size(); ++k)
        if (g_deferred[k] == me) { g_deferred.erase(g_deferred.begin() + (long)k); break; }
    if (on) g_deferred.push_back(me);
    d->deferred = on != 0;
    return 0;
}

}  // extern "C"
