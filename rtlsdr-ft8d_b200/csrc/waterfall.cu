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
//
// Execution shape (round 1c): PERSISTENT CTAs, 3 per SM.  A work item is a group of 4 consecutive frames of one slot
// (they overlap by 75 %: 1792 input samples); every CTA walks a contiguous range of groups.  Per CTA lifetime the
// twiddle/window/threshold tables are loaded into shared memory once; per group the 1024 new input samples arrive by
// cp.async into a 3-chunk ring while the previous group is being transformed, so no warp ever waits on a global load
// inside the transform.  64 threads own a frame (16 points each: stages {m=1,m=4}, {m=16,m=64}, {m=256}); the two
// exchanges go through an XOR-swizzled float2 buffer (conflict-free in all three access patterns) and are ordered by
// 64-thread named barriers, so the four frames of a CTA drift apart instead of meeting at __syncthreads.
#include "common.cuh"

namespace ft8b200 {
namespace {

constexpr int kThreads = 256;
constexpr int kFramesPerCta = 4;
constexpr int kGroupsPerSlot = kFrames / kFramesPerCta;  // 46
constexpr int kChunk = 1024;                             // input samples per ring chunk (= hop * frames per group)
constexpr int kCtasPerSm = 3;

struct cpx { float r, i; };
__device__ __forceinline__ cpx cmul(cpx a, float2 b) {  // C_MUL: each product rounded, then the add
    cpx m;
    m.r = __fsub_rn(__fmul_rn(a.r, b.x), __fmul_rn(a.i, b.y));
    m.i = __fadd_rn(__fmul_rn(a.r, b.y), __fmul_rn(a.i, b.x));
    return m;
}
__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return cpx{__fadd_rn(a.r, b.r), __fadd_rn(a.i, b.i)}; }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return cpx{__fsub_rn(a.r, b.r), __fsub_rn(a.i, b.i)}; }

// exact replacement of clamp((int)(2*(10*log10f(x))+240),0,255): count of thresholds <= x.  The MUFU.LG2 estimate is
// within one step of the answer, so one pair of independent threshold loads settles it; the loops only run if it is not.
__device__ __forceinline__ int quantise(float x, const float *__restrict__ thr) {
    int k = (int)(6.0206f * __log2f(x) + 240.0f);
    k = k < 0 ? 0 : (k > 255 ? 255 : k);
    const float lo = thr[k], hi = thr[k + 1];  // thr[0] = 0, thr[256] = +inf
    k += (x >= hi ? 1 : 0) - (x < lo ? 1 : 0);
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
__device__ __forceinline__ void bfly4_tw(cpx &f0, cpx &f1, cpx &f2, cpx &f3, float2 t1, float2 t2, float2 t3, bool trivial) {
    if (trivial) bfly4(f0, f1, f2, f3, f1, f2, f3);
    else bfly4(f0, f1, f2, f3, cmul(f1, t1), cmul(f2, t2), cmul(f3, t3));
}

// stage m=4 twiddles tw[64 q (k+1)], q = 1..3: the same for every thread -> constant bank operands
__constant__ float2 c_tw_a[3][3];

// exchange buffer index: element o of the 1024-point array lives at (o & ~15) | ((o & 15) ^ ((o >> 6) & 15)).
// Pass A writes 16 consecutive elements per thread, pass B reads/writes stride-16 and stride-64 sets, pass C reads stride 256;
// with this swizzle the 16 lanes of every half-warp hit 16 different 8-byte bank pairs in all of them.
__device__ __forceinline__ int swz(int o) { return (o & ~15) | ((o & 15) ^ ((o >> 6) & 15)); }

struct WfSmem {
    float2 tw_c[768];                  // tw[k], k < 768: last stage uses tw[i], tw[2i], tw[3i], i < 256
    float2 tw_b2[4][3][16];            // tw[4 (k+1) (i0 + 16 a)]: stage m=64
    float win[16][64];                 // window[n(j, t)]: pass A's 16 window values of thread t
    float thr[260];
    float ring_i[3][kChunk], ring_q[3][kChunk];  // input samples, chunk c of the slot in ring slot c % 3
    float2 ex[kFramesPerCta][1024];
    uint8_t out[kFramesPerCta][512];
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void frame_barrier(int fr) { asm volatile("bar.sync %0, 64;" ::"r"(fr + 1) : "memory"); }

// layout of the table blob built by build_waterfall_tables(): floats
constexpr int kBlobTwC = 0, kBlobTwB2 = kBlobTwC + 768 * 2, kBlobWin = kBlobTwB2 + 4 * 3 * 16 * 2, kBlobThr = kBlobWin + 16 * 64,
              kBlobTwB1 = kBlobThr + 260, kBlobTwA = kBlobTwB1 + 3 * 16 * 2, kBlobFloats = kBlobTwA + 9 * 2;

__global__ void __launch_bounds__(kThreads, kCtasPerSm)
waterfall1024_kernel(const float *__restrict__ d_i, const float *__restrict__ d_q, const float *__restrict__ peak,
                     const float *__restrict__ blob, int total_groups, uint8_t *__restrict__ mag) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    WfSmem &sm = *reinterpret_cast<WfSmem *>(smem_raw);
    const int tid = threadIdx.x;
    const int g_begin = (int)((long long)blockIdx.x * total_groups / gridDim.x);
    const int g_end = (int)((long long)(blockIdx.x + 1) * total_groups / gridDim.x);
    if (g_begin >= g_end) return;

    // ---- once per CTA: tables -> shared memory (contiguous in the blob in WfSmem order), stage m=16 twiddles -> registers
    {
        const float4 *src = reinterpret_cast<const float4 *>(blob);
        float4 *dst = reinterpret_cast<float4 *>(&sm.tw_c[0]);
        for (int k = tid; k < kBlobTwB1 / 4; k += kThreads) dst[k] = __ldg(src + k);
    }
    const int fr = tid >> 6, t = tid & 63;
    const int i0 = t & 15, bq = t >> 4;  // pass B coordinates
    float2 tb1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) tb1[k] = __ldg(reinterpret_cast<const float2 *>(blob + kBlobTwB1) + k * 16 + i0);
    // pass A coordinates: this thread transforms the 16 elements o = 16 p + j, which come from input n = rev2(j)*64 + r3
    // (base-4 digit reversal, kf_work's leaf copy order); p and r3 are digit reversals of each other.  The lane -> p map is
    // chosen so that both the input reads (bank = r3 mod 32) and the swizzled float2 writes are conflict-free.
    const int r3 = ((t >> 4) << 4) | ((t & 3) << 2) | ((t >> 2) & 3);
    const int p = (((t >> 2) & 3) << 4) | ((t & 3) << 2) | (t >> 4);
    float2 *ex = sm.ex[fr];

    auto load_chunk = [&](int slot, int chunk) {  // 1024 samples (the slot's last chunk, 46, holds 896), one 16-byte piece per thread and rail
        const int s = chunk * kChunk + tid * 4;
        if (s < kSlot) {
            cp_async16(&sm.ring_i[chunk % 3][tid * 4], d_i + (size_t)slot * kSlot + s);
            cp_async16(&sm.ring_q[chunk % 3][tid * 4], d_q + (size_t)slot * kSlot + s);
        }
    };

    float scale = 1.0f;
    for (int G = g_begin; G < g_end; ++G) {
        const int slot = G / kGroupsPerSlot, g = G - slot * kGroupsPerSlot;
        if (G == g_begin || g == 0) {  // (re)start of a slot: both chunks of this group are fetched now
            __syncthreads();           // the ring may still be read by the previous group
            load_chunk(slot, g);
            load_chunk(slot, g + 1);
            cp_async_commit();
            scale = 1.0f;
            if (peak != nullptr) {  // decoder(): maxSig = 0.5 / max(1e-24f, peak), rtlsdr_ft8d.c:249-259
                float pk = __ldg(peak + slot);
                if (!(pk > 1e-24f)) pk = 1e-24f;
                scale = __double2float_rn(__ddiv_rn(0.5, (double)pk));
            }
        }
        cp_async_wait_all();
        __syncthreads();  // this group's samples are visible to everyone; everyone is done with the group before
        if (g + 1 < kGroupsPerSlot && G + 1 < g_end) {  // next group's new chunk, in flight during this group's transform
            load_chunk(slot, g + 2);
            cp_async_commit();
        }
        const float *ri0 = sm.ring_i[g % 3], *rq0 = sm.ring_q[g % 3];
        const float *ri1 = sm.ring_i[(g + 1) % 3] - kChunk, *rq1 = sm.ring_q[(g + 1) % 3] - kChunk;

        cpx e[16];
        {   // pass A: stages m=1 and m=4 on 16 consecutive (digit-reversed) points
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int c = ((j & 3) << 2) | (j >> 2);
                const int s = 256 * fr + 64 * c + r3;                   // sample index inside the group's 1792-sample span
                const bool second = (4 * fr + c) >= 16;                 // warp-uniform: s >= 1024
                const float xi = second ? ri1[s] : ri0[s], xq = second ? rq1[s] : rq0[s];
                const float w = sm.win[j][t];
                e[j].r = __fmul_rn(__fmul_rn(xi, scale), w);            // decoder()'s scale, then the window (x * 1.0f is exact)
                e[j].i = __fmul_rn(__fmul_rn(xq, scale), w);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) bfly4(e[4 * q], e[4 * q + 1], e[4 * q + 2], e[4 * q + 3], e[4 * q + 1], e[4 * q + 2], e[4 * q + 3]);
            bfly4(e[0], e[4], e[8], e[12], e[4], e[8], e[12]);
#pragma unroll
            for (int q = 1; q < 4; ++q) bfly4_tw(e[q], e[q + 4], e[q + 8], e[q + 12], c_tw_a[q - 1][0], c_tw_a[q - 1][1], c_tw_a[q - 1][2], false);
#pragma unroll
            for (int j = 0; j < 16; ++j) ex[16 * p + (j ^ (t & 15))] = make_float2(e[j].r, e[j].i);  // swz(16 p + j): (o >> 6) & 15 == t & 15
        }
        frame_barrier(fr);
        {   // pass B: points 256 b + i0 + 16 a + 64 c: stages m=16 (over a) and m=64 (over c)
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const float2 v = ex[256 * bq + 64 * c + 16 * a + (i0 ^ (4 * bq + c))];
                    e[4 * c + a].r = v.x; e[4 * c + a].i = v.y;
                }
#pragma unroll
            for (int c = 0; c < 4; ++c) bfly4_tw(e[4 * c], e[4 * c + 1], e[4 * c + 2], e[4 * c + 3], tb1[0], tb1[1], tb1[2], i0 == 0);
#pragma unroll
            for (int a = 0; a < 4; ++a)
                bfly4_tw(e[a], e[4 + a], e[8 + a], e[12 + a], sm.tw_b2[a][0][i0], sm.tw_b2[a][1][i0], sm.tw_b2[a][2][i0], a == 0 && i0 == 0);
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int a = 0; a < 4; ++a) ex[256 * bq + 64 * c + 16 * a + (i0 ^ (4 * bq + c))] = make_float2(e[4 * c + a].r, e[4 * c + a].i);
        }
        frame_barrier(fr);
        {   // pass C: last stage (m = 256); only bins i and i+256 are kept (the daemon stores bins 0..511)
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = t + 64 * u;
                const int hi = (t & 48) + 64 * u, lo = t & 15;
                const float2 v0 = ex[hi + (lo ^ u)], v1 = ex[hi + 256 + (lo ^ (u + 4))], v2 = ex[hi + 512 + (lo ^ (u + 8))], v3 = ex[hi + 768 + (lo ^ (u + 12))];
                cpx f0{v0.x, v0.y}, f1{v1.x, v1.y}, f2{v2.x, v2.y}, f3{v3.x, v3.y};
                cpx a, b, c;
                if (i == 0) { a = f1; b = f2; c = f3; }
                else { a = cmul(f1, sm.tw_c[i]); b = cmul(f2, sm.tw_c[2 * i]); c = cmul(f3, sm.tw_c[3 * i]); }
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
        frame_barrier(fr);
        if (t < 32) {  // the frame's 512 bytes, 16 per lane
            uint4 *dst = reinterpret_cast<uint4 *>(mag + (size_t)slot * kWfBytes + (size_t)(g * kFramesPerCta + fr) * 512);
            dst[t] = reinterpret_cast<const uint4 *>(sm.out[fr])[t];
        }
    }
}

}  // namespace

// Host side of the table blob (layout: kBlob* above).  window/twiddles/thresholds are the host-libm tables of tables.cu.
void build_waterfall_tables(const float *window, const float2 *tw, const float *thr257, float *blob) {
    float2 *twc = reinterpret_cast<float2 *>(blob + kBlobTwC);
    for (int k = 0; k < 768; ++k) twc[k] = tw[k];
    float2 *b2 = reinterpret_cast<float2 *>(blob + kBlobTwB2);
    for (int a = 0; a < 4; ++a)
        for (int k = 0; k < 3; ++k)
            for (int i0 = 0; i0 < 16; ++i0) b2[(a * 3 + k) * 16 + i0] = tw[4 * (k + 1) * (i0 + 16 * a)];
    for (int j = 0; j < 16; ++j)
        for (int t = 0; t < 64; ++t) {
            const int c = ((j & 3) << 2) | (j >> 2);
            const int r3 = ((t >> 4) << 4) | ((t & 3) << 2) | ((t >> 2) & 3);
            blob[kBlobWin + j * 64 + t] = window[c * 64 + r3];
        }
    for (int k = 0; k < 260; ++k) blob[kBlobThr + k] = k < 257 ? thr257[k] : 0.0f;
    float2 *b1 = reinterpret_cast<float2 *>(blob + kBlobTwB1);
    for (int k = 0; k < 3; ++k)
        for (int i0 = 0; i0 < 16; ++i0) b1[k * 16 + i0] = tw[16 * (k + 1) * i0];
    float2 *ta = reinterpret_cast<float2 *>(blob + kBlobTwA);
    for (int q = 1; q < 4; ++q)
        for (int k = 0; k < 3; ++k) ta[(q - 1) * 3 + k] = tw[64 * q * (k + 1)];
}
int waterfall_blob_floats() { return kBlobFloats; }
cudaError_t upload_waterfall_constants(const float *blob_host) {
    return cudaMemcpyToSymbol(c_tw_a, blob_host + kBlobTwA, sizeof(float2) * 9);
}

cudaError_t launch_waterfall(const DeviceTables &tb, const float *d_i, const float *d_q, const float *d_peak, int n_slots, uint8_t *d_mag,
                             int sm_count, cudaStream_t st, int *launches) {
    static_assert(kFrames % kFramesPerCta == 0, "184 frames = 46 groups of 4");
    static_assert(offsetof(WfSmem, tw_b2) == kBlobTwB2 * 4 && offsetof(WfSmem, win) == kBlobWin * 4 && offsetof(WfSmem, thr) == kBlobThr * 4 &&
                      offsetof(WfSmem, ring_i) == kBlobTwB1 * 4,
                  "the blob's shared-memory part mirrors WfSmem");
    static_assert(kBlobTwB1 % 4 == 0, "blob is copied in 16-byte pieces");
    cudaError_t e = cudaFuncSetAttribute(waterfall1024_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WfSmem));
    if (e != cudaSuccess) return e;
    const int total = n_slots * kGroupsPerSlot;
    int grid = sm_count * kCtasPerSm;
    if (grid > total) grid = total;
    waterfall1024_kernel<<<grid, kThreads, sizeof(WfSmem), st>>>(d_i, d_q, d_peak, tb.wf_blob, total, d_mag);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace ft8b200
