// ptxas (CUDA 12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into ONE FFMA2 -- explicit .rn on both, -fmad=false given -- while it
// leaves the scalar mul.rn.f32 + add.rn.f32 pair alone.  A fused tap does not round the product, so the order-sensitive FIR
// (rtlsdr_ft8d.c:179-192) cannot be written with the packed mul/add pair.  Three forms that keep the two roundings:
//   ka: fma.rn.f32x2(a, b, 0) then fma.rn.f32x2(p, {1,1}, c)  -> FFMA2 (RZ addend) + FADD2      (used by cic_comb_fir_kernel)
//   kb: mul.rn.f32x2 then two scalar adds                      -> FMUL2 + 2 FADD
//   kc: two scalar muls then add.rn.f32x2                      -> 2 FMUL + FADD2
// No GPU needed:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -cubin -o /tmp/p.cubin tools/f32x2_contract_probe.cu
//                 cuobjdump -sass /tmp/p.cubin | grep -E "Function|FFMA|FMUL|FADD"
typedef unsigned long long u64;
__device__ __forceinline__ u64 mul2_rn(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2_rn(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2_rn(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__global__ void k_packed_pair_is_contracted(u64 *p) { p[3] = add2_rn(p[2], mul2_rn(p[0], p[1])); }
__global__ void k_scalar_pair_is_not(float *p) {
    float m, r;
    asm("mul.rn.f32 %0, %1, %2;" : "=f"(m) : "f"(p[0]), "f"(p[1]));
    asm("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(p[2]), "f"(m));
    p[3] = r;
}
__global__ void ka(u64 *p) { const u64 m = fma2_rn(p[0], p[1], 0ull); p[3] = fma2_rn(m, pk(1.0f, 1.0f), p[2]); }
__global__ void kb(u64 *p) {
    float m0, m1, c0, c1;
    upk(mul2_rn(p[0], p[1]), m0, m1); upk(p[2], c0, c1);
    p[3] = pk(__fadd_rn(c0, m0), __fadd_rn(c1, m1));
}
__global__ void kc(u64 *p) {
    float a0, a1, b0, b1;
    upk(p[0], a0, a1); upk(p[1], b0, b1);
    p[3] = add2_rn(p[2], pk(__fmul_rn(a0, b0), __fmul_rn(a1, b1)));
}
