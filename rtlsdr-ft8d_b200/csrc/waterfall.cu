// waterfall.cu -- windowed STFT -> uint8 dB waterfall of one 15 s slot (daemon path, 3200 sps complex).
// Replaces the FFTW loop of ft8_subsystem(), /root/reference/rtlsdr_ft8d.c:1395-1435 (+ window :331-334).
//
// 184 frames per slot (92 blocks x 2 time subdivisions), 1024-point complex FFT each, hop 256.
// The FFT is a shared-memory radix-4 decimation-in-time transform whose butterflies perform the same
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
__device__ __forceinline__ int pad(int i) { return i + (i >> 5); }

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

__global__ void __launch_bounds__(kThreads)
waterfall1024_kernel(const float *__restrict__ d_i, const float *__restrict__ d_q, const float *__restrict__ peak,
                     const float *__restrict__ window, const float2 *__restrict__ tw, const float *__restrict__ thr_g,
                     uint8_t *__restrict__ mag) {
    __shared__ float s_re[kNfft + 32], s_im[kNfft + 32];
    __shared__ float s_thr[257];
    __shared__ __align__(16) uint8_t s_out[512];
    const int frame = blockIdx.x, slot = blockIdx.y, t = threadIdx.x;
    const int start = (frame >> 1) * 512 + (frame & 1) * 256;  // idx_block*BLOCK_SIZE + time_sub*SUB_BLOCK_SIZE
    const float *xi = d_i + (size_t)slot * kSlot + start;
    const float *xq = d_q + (size_t)slot * kSlot + start;
    for (int k = t; k < 257; k += kThreads) s_thr[k] = thr_g[k];
    float scale = 1.0f;
    const bool scaled = (peak != nullptr);
    if (scaled) {  // decoder(): maxSig = 0.5 / max(1e-24f, peak), rtlsdr_ft8d.c:249-259
        float p = peak[slot];
        if (!(p > 1e-24f)) p = 1e-24f;
        scale = __double2float_rn(__ddiv_rn(0.5, (double)p));
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int n = t + kThreads * r;
        float a = xi[n], b = xq[n];
        if (scaled) { a = __fmul_rn(a, scale); b = __fmul_rn(b, scale); }
        const float w = window[n];
        // base-4 digit reversal of the 10-bit index: kiss_fft's leaf copy order (kf_work, kiss_fft.c:273-278)
        unsigned x = __brev((unsigned)n) >> 22;
        x = ((x & 0x155u) << 1) | ((x >> 1) & 0x155u);
        s_re[pad((int)x)] = __fmul_rn(a, w);
        s_im[pad((int)x)] = __fmul_rn(b, w);
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int m = 1 << (2 * s);
        const int i = t & (m - 1);
        const int base = ((t >> (2 * s)) << (2 * s + 2)) + i;
        const int fs = 256 >> (2 * s);
        const int p0 = pad(base), p1 = pad(base + m), p2 = pad(base + 2 * m), p3 = pad(base + 3 * m);
        cpx f0{s_re[p0], s_im[p0]}, f1{s_re[p1], s_im[p1]}, f2{s_re[p2], s_im[p2]}, f3{s_re[p3], s_im[p3]};
        const cpx a = cmul(f1, __ldg(&tw[i * fs]));
        const cpx b = cmul(f2, __ldg(&tw[2 * i * fs]));
        const cpx c = cmul(f3, __ldg(&tw[3 * i * fs]));
        const cpx d5 = csub(f0, b);
        f0 = cadd(f0, b);
        const cpx s3 = cadd(a, c);
        const cpx s4 = csub(a, c);
        f2 = csub(f0, s3);
        f0 = cadd(f0, s3);
        f1.r = __fadd_rn(d5.r, s4.i); f1.i = __fsub_rn(d5.i, s4.r);
        f3.r = __fsub_rn(d5.r, s4.i); f3.i = __fadd_rn(d5.i, s4.r);
        s_re[p0] = f0.r; s_im[p0] = f0.i;
        s_re[p1] = f1.r; s_im[p1] = f1.i;
        s_re[p2] = f2.r; s_im[p2] = f2.i;
        s_re[p3] = f3.r; s_im[p3] = f3.i;
        __syncthreads();
    }
    {   // last pass (m = 256): only bins t and t+256 are needed (the daemon keeps bins 0..511)
        const int p0 = pad(t), p1 = pad(t + 256), p2 = pad(t + 512), p3 = pad(t + 768);
        cpx f0{s_re[p0], s_im[p0]}, f1{s_re[p1], s_im[p1]}, f2{s_re[p2], s_im[p2]}, f3{s_re[p3], s_im[p3]};
        const cpx a = cmul(f1, __ldg(&tw[t]));
        const cpx b = cmul(f2, __ldg(&tw[2 * t]));
        const cpx c = cmul(f3, __ldg(&tw[3 * t]));
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
        s_out[(t & 1) * 256 + (t >> 1)] = (uint8_t)quantise(x0, s_thr);
        s_out[(t & 1) * 256 + 128 + (t >> 1)] = (uint8_t)quantise(x1, s_thr);
    }
    __syncthreads();
    if (t < 32) {
        uint4 *dst = reinterpret_cast<uint4 *>(mag + (size_t)slot * kWfBytes + (size_t)frame * 512);
        dst[t] = reinterpret_cast<const uint4 *>(s_out)[t];
    }
}

}  // namespace

cudaError_t launch_waterfall(const DeviceTables &tb, const float *d_i, const float *d_q, const float *d_peak, int n_slots, uint8_t *d_mag,
                             cudaStream_t st, int *launches) {
    dim3 grid(kFrames, n_slots);
    waterfall1024_kernel<<<grid, kThreads, 0, st>>>(d_i, d_q, d_peak, tb.window1024, tb.twiddle1024, tb.db_thresholds, d_mag);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace ft8b200
