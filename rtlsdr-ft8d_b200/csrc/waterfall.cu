// waterfall.cu -- windowed STFT -> uint8 dB waterfall of one 15 s slot (daemon path, 3200 sps complex).
// Replaces the FFTW loop of ft8_subsystem(), /root/reference/rtlsdr_ft8d.c:1395-1435 (+ window :331-334).
//
// 184 frames per slot (92 blocks x 2 time subdivisions), 1024-point complex FFT each, hop 256.
// The FFT is a register/shared-memory radix-4 decimation-in-time transform whose butterflies perform the same
// float operations, in the same order, as the reference's vendored kiss_fft (kf_bfly4,
// ft8_lib/fft/kiss_fft.c:38-84 with C_MUL of _kiss_fft_guts.h:81-83) -- 1024 = 4^5, so kiss_fft uses
// five radix-4 passes over a base-4 digit-reversed input -- which makes the spectrum, and therefore
// every waterfall byte, bit-identical to the CPU path that uses kiss_fft.  (The reference daemon itself
// links FFTW3f, an external library whose rounding depends on its plan; see DESIGN.md.)
// Magnitude -> dB -> uint8 is fused into the last pass; log10f is replaced by a comparison against the
// 255 host-computed step thresholds of the reference's quantiser (tables.cu), which is exact.
#include "common.cuh"

namespace ft8b200 {
namespace {

constexpr int kThreads = 256;

struct cpx { float r, i; };
__device__ __forceinline__ cpx cmul(cpx a, float2 b) {  // C_MUL: each product rounded, then the add
    cpx m;
    m.r = __fsub_rn(__fmul_rn(a.r, b.x), __fmul_rn(a.i, b.y));
    m.i = __fadd_rn(__fmul_rn(a.r, b.y), __fmul_rn(a.i, b.x));
    return m;
}
__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return cpx{__fadd_rn(a.r, b.r), __fadd_rn(a.i, b.i)}; }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return cpx{__fsub_rn(a.r, b.r), __fsub_rn(a.i, b.i)}; }

// exact replacement of clamp((int)(2*(10*log10f(x))+240),0,255): count of thresholds <= x
__device__ __forceinline__ int quantise(float x, const float *__restrict__ thr) {
    int k = (int)(6.0206f * __log2f(x) + 240.0f);
    k = k < 0 ? 0 : (k > 255 ? 255 : k);
    while (k > 0 && x < thr[k]) --k;
    while (k < 255 && x >= thr[k + 1]) ++k;
    return k;
}

// radix-4 butterfly of kf_bfly4 (forward), twiddled inputs a,b,c already formed
__device__ __forceinline__ void bfly4(cpx &f0, cpx &f1, cpx &f2, cpx &f3, const cpx a, const cpx b, const cpx c) {
    const cpx d5 = csub(f0, b);
    f0 = cadd(f0, b);
    const cpx s3 = cadd(a, c);
    const cpx s4 = csub(a, c);
    f2 = csub(f0, s3);
    f0 = cadd(f0, s3);
    f1.r = __fadd_rn(d5.r, s4.i); f1.i = __fsub_rn(d5.i, s4.r);
    f3.r = __fsub_rn(d5.r, s4.i); f3.i = __fadd_rn(d5.i, s4.r);
}
// kiss_fft multiplies by twiddle 0 = (1, -0) like by any other; the product equals the input except possibly for the
// sign of a zero, which cannot reach |X|^2 -- so index-0 twiddles are skipped.
__device__ __forceinline__ void bfly4_tw(cpx &f0, cpx &f1, cpx &f2, cpx &f3, const float2 *__restrict__ tw, int i1, int i2, int i3, bool trivial) {
    if (trivial) bfly4(f0, f1, f2, f3, f1, f2, f3);
    else bfly4(f0, f1, f2, f3, cmul(f1, __ldg(&tw[i1])), cmul(f2, __ldg(&tw[i2])), cmul(f3, __ldg(&tw[i3])));
}

constexpr int kFramesPerCta = 4;
constexpr int kFrameThreads = 64;
constexpr int kSpan = 1024 + 256 * (kFramesPerCta - 1);  // input samples covered by the CTA's frames
constexpr int kPadLen = 1024 + 64;                       // padded FFT array: index o -> o + (o >> 4)
__device__ __forceinline__ int pad16(int o) { return o + (o >> 4); }

struct WfSmem {
    float xi[kSpan], xq[kSpan];                       // scaled input samples
    float re[kFramesPerCta][kPadLen], im[kFramesPerCta][kPadLen];
    float thr[260];
    uint8_t out[kFramesPerCta][512];
};

// One CTA = 4 consecutive frames (they overlap by 75 %, so their 1792 input samples are staged once);
// 64 threads per frame, each holding 16 points: passes {m=1,m=4}, {m=16,m=64}, {m=256} of the five radix-4 stages.
__global__ void __launch_bounds__(kThreads)
waterfall1024_kernel(const float *__restrict__ d_i, const float *__restrict__ d_q, const float *__restrict__ peak,
                     const float *__restrict__ window, const float2 *__restrict__ tw, const float *__restrict__ thr_g,
                     uint8_t *__restrict__ mag) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    WfSmem &sm = *reinterpret_cast<WfSmem *>(smem_raw);
    const int tid = threadIdx.x, slot = blockIdx.y;
    const int frame0 = blockIdx.x * kFramesPerCta;
    const int start0 = (frame0 >> 1) * 512 + (frame0 & 1) * 256;  // == 256 * frame0
    const float *xi = d_i + (size_t)slot * kSlot + start0;
    const float *xq = d_q + (size_t)slot * kSlot + start0;
    for (int k = tid; k < 257; k += kThreads) sm.thr[k] = thr_g[k];
    float scale = 1.0f;
    const bool scaled = (peak != nullptr);
    if (scaled) {  // decoder(): maxSig = 0.5 / max(1e-24f, peak), rtlsdr_ft8d.c:249-259
        float p = peak[slot];
        if (!(p > 1e-24f)) p = 1e-24f;
        scale = __double2float_rn(__ddiv_rn(0.5, (double)p));
    }
    for (int k = tid; k < kSpan; k += kThreads) {
        float a = xi[k], b = xq[k];
        if (scaled) { a = __fmul_rn(a, scale); b = __fmul_rn(b, scale); }
        sm.xi[k] = a;
        sm.xq[k] = b;
    }
    __syncthreads();

    const int fr = tid >> 6, t = tid & 63;
    float *re = sm.re[fr], *im = sm.im[fr];
    cpx e[16];
    {   // pass A: 16 consecutive (digit-reversed) points: stages m=1 and m=4
        // element o = 16 t + j comes from input n = rev2(j) * 64 + rev3(t) (base-4 digit reversal, kf_work's leaf copy order)
        const int r3 = ((t & 3) << 4) | (t & 12) | (t >> 4);
        const float *fi = sm.xi + 256 * fr, *fq = sm.xq + 256 * fr;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int n = (((j & 3) << 2) | (j >> 2)) * 64 + r3;
            const float w = __ldg(&window[n]);
            e[j].r = __fmul_rn(fi[n], w);
            e[j].i = __fmul_rn(fq[n], w);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) bfly4(e[4 * q], e[4 * q + 1], e[4 * q + 2], e[4 * q + 3], e[4 * q + 1], e[4 * q + 2], e[4 * q + 3]);
#pragma unroll
        for (int q = 0; q < 4; ++q) bfly4_tw(e[q], e[q + 4], e[q + 8], e[q + 12], tw, 64 * q, 128 * q, 192 * q, q == 0);
#pragma unroll
        for (int j = 0; j < 16; ++j) { re[pad16(16 * t + j)] = e[j].r; im[pad16(16 * t + j)] = e[j].i; }
    }
    __syncthreads();
    {   // pass B: points 256 b + i0 + 16 a + 64 c: stages m=16 (over a) and m=64 (over c)
        const int i0 = t & 15, b = t >> 4;
        const int base = 256 * b + i0;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int a = 0; a < 4; ++a) { const int o = pad16(base + 16 * a + 64 * c); e[4 * c + a].r = re[o]; e[4 * c + a].i = im[o]; }
#pragma unroll
        for (int c = 0; c < 4; ++c) bfly4_tw(e[4 * c], e[4 * c + 1], e[4 * c + 2], e[4 * c + 3], tw, 16 * i0, 32 * i0, 48 * i0, i0 == 0);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int i = i0 + 16 * a;
            bfly4_tw(e[a], e[4 + a], e[8 + a], e[12 + a], tw, 4 * i, 8 * i, 12 * i, i == 0);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int a = 0; a < 4; ++a) { const int o = pad16(base + 16 * a + 64 * c); re[o] = e[4 * c + a].r; im[o] = e[4 * c + a].i; }
    }
    __syncthreads();
    {   // pass C: last stage (m = 256); only bins i and i+256 are kept (the daemon stores bins 0..511)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = t + 64 * u;
            cpx f0{re[pad16(i)], im[pad16(i)]}, f1{re[pad16(i + 256)], im[pad16(i + 256)]};
            cpx f2{re[pad16(i + 512)], im[pad16(i + 512)]}, f3{re[pad16(i + 768)], im[pad16(i + 768)]};
            cpx a, b, c;
            if (i == 0) { a = f1; b = f2; c = f3; }
            else { a = cmul(f1, __ldg(&tw[i])); b = cmul(f2, __ldg(&tw[2 * i])); c = cmul(f3, __ldg(&tw[3 * i])); }
            const cpx d5 = csub(f0, b);
            f0 = cadd(f0, b);
            const cpx s3 = cadd(a, c);
            const cpx s4 = csub(a, c);
            f0 = cadd(f0, s3);
            f1.r = __fadd_rn(d5.r, s4.i); f1.i = __fsub_rn(d5.i, s4.r);
            // mag2 * 4.0f / (NFFT*NFFT) then 1E-12f + ..., rtlsdr_ft8d.c:1415-1416 (the divide by 2^20 is an exact scaling)
            const float m0 = __fadd_rn(__fmul_rn(f0.r, f0.r), __fmul_rn(f0.i, f0.i));
            const float m1 = __fadd_rn(__fmul_rn(f1.r, f1.r), __fmul_rn(f1.i, f1.i));
            const float x0 = __fadd_rn(1E-12f, __fmul_rn(__fmul_rn(m0, 4.0f), 9.5367431640625e-07f));
            const float x1 = __fadd_rn(1E-12f, __fmul_rn(__fmul_rn(m1, 4.0f), 9.5367431640625e-07f));
            // layout [freq_sub][bin]: FFT bin 2*bin+freq_sub, rtlsdr_ft8d.c:1420-1428
            sm.out[fr][(i & 1) * 256 + (i >> 1)] = (uint8_t)quantise(x0, sm.thr);
            sm.out[fr][(i & 1) * 256 + 128 + (i >> 1)] = (uint8_t)quantise(x1, sm.thr);
        }
    }
    __syncthreads();
    if (tid < kFramesPerCta * 32) {
        uint4 *dst = reinterpret_cast<uint4 *>(mag + (size_t)slot * kWfBytes + (size_t)frame0 * 512);
        dst[tid] = reinterpret_cast<const uint4 *>(&sm.out[0][0])[tid];
    }
}

}  // namespace

cudaError_t launch_waterfall(const DeviceTables &tb, const float *d_i, const float *d_q, const float *d_peak, int n_slots, uint8_t *d_mag,
                             cudaStream_t st, int *launches) {
    static_assert(kFrames % kFramesPerCta == 0, "184 frames = 46 CTAs of 4");
    cudaError_t e = cudaFuncSetAttribute(waterfall1024_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WfSmem));
    if (e != cudaSuccess) return e;
    dim3 grid(kFrames / kFramesPerCta, n_slots);
    waterfall1024_kernel<<<grid, kThreads, sizeof(WfSmem), st>>>(d_i, d_q, d_peak, tb.window1024, tb.twiddle1024, tb.db_thresholds, d_mag);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace ft8b200
