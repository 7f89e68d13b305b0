// Does a packed FP32 instruction (FADD2, sm_100) cost one issue slot or two?  Per SM sub-partition the FP32 pipe finishes one scalar warp
// instruction per cycle, one packed per two cycles (tools/f32x2_probe.cu).  If a packed instruction took ONE issue slot, integer ALU
// instructions could issue in the second cycle for free: N FADD2 + N integer adds would take 2N cycles, like N FADD2 alone.  If it
// holds the scheduler for both cycles, the mix takes 3N -- exactly what 2N scalar FADD + N integer adds take.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/build/f32x2_issue_probe tools/f32x2_issue_probe.cu
// Prints ms and cycles per loop trip per scheduler for: (0) 8 FADD2, (1) 8 IADD, (2) 8 FADD2 + 8 IADD, (3) 16 FADD, (4) 16 FADD + 8 IADD.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define ADD2(x) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(s2))
#define ADD1(x) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x) : "f"(s1))
#define IADD(x) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(si))
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float s1, unsigned si) {
    u64 s2; asm("mov.b64 %0, {%1,%1};" : "=l"(s2) : "f"(s1));
    u64 p0 = threadIdx.x, p1 = p0 + 1, p2 = p0 + 2, p3 = p0 + 3, p4 = p0 + 4, p5 = p0 + 5, p6 = p0 + 6, p7 = p0 + 7;
    float f[16];
    unsigned i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3, i4 = i0 + 4, i5 = i0 + 5, i6 = i0 + 6, i7 = i0 + 7;
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = threadIdx.x + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (MODE == 0 || MODE == 2) { ADD2(p0); ADD2(p1); ADD2(p2); ADD2(p3); ADD2(p4); ADD2(p5); ADD2(p6); ADD2(p7); }
            if (MODE == 3 || MODE == 4) {
#pragma unroll
                for (int j = 0; j < 16; ++j) ADD1(f[j]);
            }
            if (MODE == 1 || MODE == 2 || MODE == 4) { IADD(i0); IADD(i1); IADD(i2); IADD(i3); IADD(i4); IADD(i5); IADD(i6); IADD(i7); }
        }
    }
    float acc = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) acc += f[j];
    u64 q = p0 ^ p1 ^ p2 ^ p3 ^ p4 ^ p5 ^ p6 ^ p7;
    unsigned r = i0 ^ i1 ^ i2 ^ i3 ^ i4 ^ i5 ^ i6 ^ i7;
    if (acc == 12345.678f || q == 42 || r == 43) out[0] = acc + (float)q + (float)r;
}
template <int MODE> void run(const char *name, float *d, int sms, int mhz) {
    const int iters = 20000, ctas = sms * 2;   // 2 CTAs x 8 warps per SM = 4 warps per scheduler
    k<MODE><<<ctas, 256>>>(d, 100, 1.0f, 3); cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<MODE><<<ctas, 256>>>(d, iters, 1.0f, 3); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    // per scheduler: 4 warps x iters x 4 unrolled trips
    const double cyc = ms * 1e-3 * mhz * 1e6 / (4.0 * iters * 4);
    printf("mode %d %-28s %8.3f ms   %6.2f cycles per warp trip per scheduler (at %d MHz)\n", MODE, name, ms, cyc, mhz);
}
int main() {
    float *d; cudaMalloc(&d, 4);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int mhz = 0; cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0); mhz /= 1000;
    run<0>("8 FADD2", d, p.multiProcessorCount, mhz);
    run<1>("8 IADD", d, p.multiProcessorCount, mhz);
    run<2>("8 FADD2 + 8 IADD", d, p.multiProcessorCount, mhz);
    run<3>("16 FADD", d, p.multiProcessorCount, mhz);
    run<4>("16 FADD + 8 IADD", d, p.multiProcessorCount, mhz);
    return 0;
}
