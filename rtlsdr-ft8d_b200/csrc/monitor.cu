// monitor.cu -- ft8_lib's "monitor": real audio (12 kHz) -> Hann-windowed STFT -> uint8 dB waterfall.
// Replaces monitor_init/process/reset/free and the kiss_fftr loop, /root/reference/ft8_lib/decode_ft8.c:35-39,
// 111-224, ft8_lib/fft/kiss_fftr.c:22-115, ft8_lib/fft/kiss_fft.c:15-382.
//
// One CTA per analysis frame.  A frame of nfft real samples (3840 for FT8 at 12 kHz) is transformed the way
// kiss_fftr does it -- as an nfft/2-point complex FFT of the even/odd-packed signal followed by the split
// step with "super twiddles" -- and the complex FFT follows kiss_fft's own decomposition (radix 4 first, then
// 2, 3, 5; 1920 = 4*4*4*2*3*5) with its butterflies' exact operation order, so spectra and waterfall bytes are
// bit-identical to the CPU path.  The stage list, the input permutation and all twiddles are built on the host
// (same libm calls as kiss_fft) and read from global memory; data lives in shared memory between stages.
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

__device__ __forceinline__ int quantise(float x, const float *__restrict__ thr) {
    int k = (int)(6.0206f * __log2f(x) + 240.0f);
    k = k < 0 ? 0 : (k > 255 ? 255 : k);
    while (k > 0 && x < thr[k]) --k;
    while (k < 255 && x >= thr[k + 1]) ++k;
    return k;
}

// dynamic smem: re[n], im[n], thr[257]
__global__ void __launch_bounds__(kMonThreads)
monitor_frames_kernel(const float *__restrict__ audio, size_t slot_stride, int n_samples, long first_start, int hop, int nfft, FftPlan plan,
                      const uint16_t *__restrict__ perm, const float2 *__restrict__ tw, const float2 *__restrict__ super_tw,
                      const float *__restrict__ wnorm, const float *__restrict__ thr_g, int num_bins, int freq_osr, uint8_t *__restrict__ mag,
                      size_t mag_slot_stride, unsigned int *__restrict__ xmax_bits) {
    extern __shared__ float smem[];
    const int n = plan.n;
    float *re = smem, *im = smem + n, *thr = smem + 2 * n;
    const int t = threadIdx.x, frame = blockIdx.x, slot = blockIdx.y;
    const float *x = audio + (size_t)slot * slot_stride;
    const long start = first_start + (long)frame * hop;
    for (int k = t; k < 257; k += kMonThreads) thr[k] = thr_g[k];
    // windowed, normalised frame, packed (even, odd) -> complex, stored in kiss_fft's permuted leaf order
    for (int o = t; o < n; o += kMonThreads) {
        const int src = perm[o];
        const long p0 = start + 2 * src, p1 = p0 + 1;
        const float a = (p0 >= 0 && p0 < n_samples) ? x[p0] : 0.0f;
        const float b = (p1 >= 0 && p1 < n_samples) ? x[p1] : 0.0f;
        re[o] = __fmul_rn(wnorm[2 * src], a);      // (fft_norm * window[pos]) * last_frame[pos], decode_ft8.c:191
        im[o] = __fmul_rn(wnorm[2 * src + 1], b);
    }
    __syncthreads();
    for (int s = 0; s < plan.nstages; ++s) {
        const int p = plan.radix[s], m = plan.m[s], fs = plan.fstride[s];
        const int nbf = n / p;  // butterflies in this stage
        for (int b = t; b < nbf; b += kMonThreads) {
            const int g = b / m, k = b - g * m;
            const int i0 = g * p * m + k;
            if (p == 4) {  // kf_bfly4, kiss_fft.c:38-84
                cpx f0{re[i0], im[i0]}, f1{re[i0 + m], im[i0 + m]}, f2{re[i0 + 2 * m], im[i0 + 2 * m]}, f3{re[i0 + 3 * m], im[i0 + 3 * m]};
                const cpx a = cmul(f1, tw[k * fs]);
                const cpx bb = cmul(f2, tw[2 * k * fs]);
                const cpx c = cmul(f3, tw[3 * k * fs]);
                const cpx d5 = csub(f0, bb);
                f0 = cadd(f0, bb);
                const cpx s3 = cadd(a, c), s4 = csub(a, c);
                f2 = csub(f0, s3);
                f0 = cadd(f0, s3);
                re[i0] = f0.r; im[i0] = f0.i;
                re[i0 + 2 * m] = f2.r; im[i0 + 2 * m] = f2.i;
                re[i0 + m] = __fadd_rn(d5.r, s4.i); im[i0 + m] = __fsub_rn(d5.i, s4.r);
                re[i0 + 3 * m] = __fsub_rn(d5.r, s4.i); im[i0 + 3 * m] = __fadd_rn(d5.i, s4.r);
            } else if (p == 2) {  // kf_bfly2, kiss_fft.c:15-36
                cpx f0{re[i0], im[i0]}, f1{re[i0 + m], im[i0 + m]};
                const cpx tt = cmul(f1, tw[k * fs]);
                f1 = csub(f0, tt);
                f0 = cadd(f0, tt);
                re[i0] = f0.r; im[i0] = f0.i;
                re[i0 + m] = f1.r; im[i0 + m] = f1.i;
            } else if (p == 3) {  // kf_bfly3, kiss_fft.c:86-128
                const float2 e3 = tw[fs * m];
                cpx f0{re[i0], im[i0]}, f1{re[i0 + m], im[i0 + m]}, f2{re[i0 + 2 * m], im[i0 + 2 * m]};
                const cpx s1 = cmul(f1, tw[k * fs]);
                const cpx s2 = cmul(f2, tw[2 * k * fs]);
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
                re[i0] = f0.r; im[i0] = f0.i;
                re[i0 + m] = f1.r; im[i0 + m] = f1.i;
                re[i0 + 2 * m] = f2.r; im[i0 + 2 * m] = f2.i;
            } else {  // p == 5: kf_bfly5, kiss_fft.c:130-190
                const float2 ya = tw[fs * m], yb = tw[fs * 2 * m];
                const cpx s0{re[i0], im[i0]};
                cpx f1{re[i0 + m], im[i0 + m]}, f2{re[i0 + 2 * m], im[i0 + 2 * m]}, f3{re[i0 + 3 * m], im[i0 + 3 * m]}, f4{re[i0 + 4 * m], im[i0 + 4 * m]};
                const cpx s1 = cmul(f1, tw[k * fs]);
                const cpx s2 = cmul(f2, tw[2 * k * fs]);
                const cpx s3 = cmul(f3, tw[3 * k * fs]);
                const cpx s4 = cmul(f4, tw[4 * k * fs]);
                const cpx s7 = cadd(s1, s4), s10 = csub(s1, s4), s8 = cadd(s2, s3), s9 = csub(s2, s3);
                re[i0] = __fadd_rn(s0.r, __fadd_rn(s7.r, s8.r));
                im[i0] = __fadd_rn(s0.i, __fadd_rn(s7.i, s8.i));
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
                re[i0 + m] = f1.r; im[i0 + m] = f1.i;
                re[i0 + 2 * m] = f2.r; im[i0 + 2 * m] = f2.i;
                re[i0 + 3 * m] = f3.r; im[i0 + 3 * m] = f3.i;
                re[i0 + 4 * m] = f4.r; im[i0 + 4 * m] = f4.i;
            }
        }
        __syncthreads();
    }
    // real-FFT split (kiss_fftr.c:80-113) fused with |X|^2 -> dB -> uint8 (decode_ft8.c:196-213)
    const int wanted = num_bins * freq_osr;  // bins 0 .. wanted-1 are stored (wanted <= n)
    uint8_t *out = mag + (size_t)slot * mag_slot_stride + (size_t)frame * wanted;
    float xmax = 0.0f;
    for (int k = t; k <= n / 2; k += kMonThreads) {
        cpx lo, hi;  // bins k and n-k
        bool have_hi = false;
        if (k == 0) {
            lo.r = __fadd_rn(re[0], im[0]);
            lo.i = 0.0f;
        } else {
            const cpx fpk{re[k], im[k]};
            const cpx fpnk{re[n - k], -im[n - k]};
            const cpx f1 = cadd(fpk, fpnk), f2 = csub(fpk, fpnk);
            const cpx tt = cmul(f2, super_tw[k - 1]);
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
            out[(k % freq_osr) * num_bins + k / freq_osr] = (uint8_t)quantise(xv, thr);
        }
        const int kh = n - k;
        if (have_hi && kh != k && kh < wanted) {
            const float m2 = __fadd_rn(__fmul_rn(hi.i, hi.i), __fmul_rn(hi.r, hi.r));
            const float xv = __fadd_rn(1E-12f, m2);
            xmax = fmaxf(xmax, xv);
            out[(kh % freq_osr) * num_bins + kh / freq_osr] = (uint8_t)quantise(xv, thr);
        }
    }
    if (xmax_bits) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        if ((t & 31) == 0) atomicMax(xmax_bits + slot, __float_as_uint(xmax));
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
        if (nf >= kMaxStages || (q != 2 && q != 3 && q != 4 && q != 5)) return false;
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
    bool ok = cudaMalloc(&t->d_perm, sizeof(uint16_t) * n) == cudaSuccess && cudaMalloc(&t->d_tw, sizeof(float2) * n) == cudaSuccess &&
              cudaMalloc(&t->d_super, sizeof(float2) * (n / 2 + 1)) == cudaSuccess && cudaMalloc(&t->d_wnorm, sizeof(float) * nfft) == cudaSuccess &&
              cudaMemcpy(t->d_perm, perm.data(), sizeof(uint16_t) * n, cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(t->d_tw, tw.data(), sizeof(float2) * n, cudaMemcpyHostToDevice) == cudaSuccess &&
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
    const size_t smem = sizeof(float) * (2 * (size_t)t->plan.n + 257);
    cudaError_t e = cudaFuncSetAttribute(monitor_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid(n_frames, n_slots);
    monitor_frames_kernel<<<grid, kMonThreads, smem, st>>>(d_audio, slot_stride, n_samples, first_start, hop, t->nfft, t->plan, t->d_perm, t->d_tw,
                                                           t->d_super, t->d_wnorm, thresholds(t->device), num_bins, freq_osr, d_mag, mag_slot_stride, d_xmax);
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
    float *d_buf;          // nfft - subblock + block_size floats: history + the new block
    uint8_t *d_mag;        // one block's worth of bytes
    unsigned int *d_xmax;
    float *h_buf;          // pinned
    cudaStream_t st;
};

void die(const char *what) {
    fprintf(stderr, "libft8b200: monitor: %s (%s)\n", what, cudaGetErrorString(cudaGetLastError()));
    abort();
}

}  // namespace

extern "C" {

int ft8b200_monitor_waterfall(ft8b200_ctx_t *ctx, const float *d_audio, size_t slot_stride_samples, int n_samples, int n_slots, int sample_rate,
                              int time_osr, int freq_osr, int protocol, uint8_t *d_mag, size_t mag_slot_stride, int *num_blocks_out, void *stream) {
    if (!ctx || !d_audio || !d_mag || n_slots < 1 || n_samples < 0 || time_osr < 1 || freq_osr < 1) return FT8B200_EINVAL;
    const MonGeom g = geometry(sample_rate, time_osr, freq_osr, protocol);
    if (g.block_size < 2 || g.subblock_size * time_osr != g.block_size) return FT8B200_EINVAL;
    const int device = ctx_device(ctx);
    if (cudaSetDevice(device) != cudaSuccess) return FT8B200_ECUDA;  // tables and launches belong to the context's device, not the caller's current one
    const MonTables *t = get_tables(device, g.nfft);
    if (!t || !thresholds(device)) return FT8B200_EINVAL;
    int nb = n_samples / g.block_size;
    if (nb > g.max_blocks) nb = g.max_blocks;
    if (num_blocks_out) *num_blocks_out = nb;
    if (nb == 0) return 0;
    const size_t stride = (size_t)time_osr * freq_osr * g.num_bins;
    if (mag_slot_stride < (size_t)nb * stride) return FT8B200_EINVAL;
    cudaStream_t st = stream ? (cudaStream_t)stream : (cudaStream_t)ft8b200_cuda_stream(ctx);
    // frame f ends at sample (f+1)*subblock; the reference's last_frame starts out as (zeroed) history
    cudaError_t e = launch_frames(t, d_audio, slot_stride_samples, n_samples, (long)g.subblock_size - g.nfft, g.subblock_size, nb * time_osr, n_slots,
                                  g.num_bins, freq_osr, d_mag, mag_slot_stride, nullptr, st);
    return e == cudaSuccess ? 0 : FT8B200_ECUDA;
}

void monitor_init(monitor_t *me, const monitor_config_t *cfg) {
    ft8b200_ctx_t *ctx = default_ctx();
    const MonGeom g = geometry(cfg->sample_rate, cfg->time_osr, cfg->freq_osr, (int)cfg->protocol);
    const int device = ctx_device(ctx);
    if (cudaSetDevice(device) != cudaSuccess) die("monitor_init: cudaSetDevice");
    const MonTables *t = get_tables(device, g.nfft);
    if (!t || !thresholds(device)) {
        fprintf(stderr, "libft8b200: monitor_init: unsupported FFT size %d (radices 2,3,4,5 only)\n", g.nfft);
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
    const size_t nbuf = (size_t)g.nfft - g.subblock_size + g.block_size;
    if (cudaMalloc(&d->d_buf, sizeof(float) * nbuf) != cudaSuccess || cudaMalloc(&d->d_mag, (size_t)me->wf.block_stride) != cudaSuccess ||
        cudaMalloc(&d->d_xmax, sizeof(unsigned int)) != cudaSuccess || cudaHostAlloc(&d->h_buf, sizeof(float) * nbuf, cudaHostAllocDefault) != cudaSuccess)
        die("monitor_init: allocation failed");
    me->fft_work = d;
    me->fft_cfg = nullptr;
}

void monitor_free(monitor_t *me) {
    MonitorDev *d = (MonitorDev *)me->fft_work;
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
    me->wf.num_blocks = 0;
    me->max_mag = 0;
}

void monitor_process(monitor_t *me, const float *frame) {
    if (me->wf.num_blocks >= me->wf.max_blocks) return;  // silent no-op once full (decode_ft8.c:165-166)
    MonitorDev *d = (MonitorDev *)me->fft_work;
    if (cudaSetDevice(d->tables->device) != cudaSuccess) die("monitor_process: cudaSetDevice");  // the caller's thread may have another device current
    const int nfft = me->nfft, sub = me->subblock_size, blk = me->block_size, tosr = me->wf.time_osr;
    const size_t nhist = (size_t)nfft - sub, nbuf = nhist + blk;
    memcpy(d->h_buf, me->last_frame + sub, sizeof(float) * nhist);
    memcpy(d->h_buf + nhist, frame, sizeof(float) * (size_t)blk);
    const size_t off = (size_t)me->wf.num_blocks * me->wf.block_stride;
    unsigned int xbits = 0;
    bool ok = cudaMemcpyAsync(d->d_buf, d->h_buf, sizeof(float) * nbuf, cudaMemcpyHostToDevice, d->st) == cudaSuccess &&
              cudaMemsetAsync(d->d_xmax, 0, sizeof(unsigned int), d->st) == cudaSuccess &&
              launch_frames(d->tables, d->d_buf, 0, (int)nbuf, 0, sub, tosr, 1, me->wf.num_bins, me->wf.freq_osr, d->d_mag, 0, d->d_xmax, d->st) == cudaSuccess &&
              cudaMemcpyAsync(me->wf.mag + off, d->d_mag, (size_t)me->wf.block_stride, cudaMemcpyDeviceToHost, d->st) == cudaSuccess &&
              cudaMemcpyAsync(&xbits, d->d_xmax, sizeof(xbits), cudaMemcpyDeviceToHost, d->st) == cudaSuccess && cudaStreamSynchronize(d->st) == cudaSuccess;
    if (!ok) die("monitor_process failed");
    // host copy of the sliding frame, as the reference keeps it (decode_ft8.c:178-186)
    memcpy(me->last_frame, d->h_buf + (size_t)(tosr - 1) * sub, sizeof(float) * (size_t)nfft);
    float xmax;
    memcpy(&xmax, &xbits, 4);
    const float db = 10.0f * log10f(xmax);  // max over the block of 10*log10f(x): log10f is monotone, so this is the block's max dB
    if (db > me->max_mag) me->max_mag = db;
    ++me->wf.num_blocks;
}

}  // extern "C"
