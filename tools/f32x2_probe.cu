// Microbenchmark: scalar FADD/FMUL against the packed add/mul.rn.f32x2 (FADD2/FMUL2, sm_100+) -- lane operations per second over the whole GPU.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o /tmp/f32x2_probe tools/f32x2_probe.cu && /tmp/f32x2_probe
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ unsigned long long pk(float a, float b){ unsigned long long r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void upk(unsigned long long v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v)); }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b){ unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b){ unsigned long long r; asm("sub.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b){ unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r;}
template<int MODE>
__global__ void k(float* out, int iters, float s){
  float a0=threadIdx.x, a1=a0+1, a2=a0+2, a3=a0+3, a4=a0+4,a5=a0+5,a6=a0+6,a7=a0+7;
  if (MODE==0){
    for(int i=0;i<iters;++i){
      #pragma unroll
      for(int u=0;u<8;++u){ a0=__fadd_rn(a0,s);a1=__fadd_rn(a1,s);a2=__fadd_rn(a2,s);a3=__fadd_rn(a3,s);a4=__fadd_rn(a4,s);a5=__fadd_rn(a5,s);a6=__fadd_rn(a6,s);a7=__fadd_rn(a7,s);}
    }
  } else if (MODE==1){
    unsigned long long p0=pk(a0,a1),p1=pk(a2,a3),p2=pk(a4,a5),p3=pk(a6,a7), ss=pk(s,s);
    for(int i=0;i<iters;++i){
      #pragma unroll
      for(int u=0;u<8;++u){ p0=add2(p0,ss);p1=add2(p1,ss);p2=add2(p2,ss);p3=add2(p3,ss);}
    }
    upk(p0,a0,a1);upk(p1,a2,a3);upk(p2,a4,a5);upk(p3,a6,a7);
  } else if (MODE==2){
    for(int i=0;i<iters;++i){
      #pragma unroll
      for(int u=0;u<8;++u){ a0=__fmul_rn(a0,s);a1=__fmul_rn(a1,s);a2=__fmul_rn(a2,s);a3=__fmul_rn(a3,s);a4=__fmul_rn(a4,s);a5=__fmul_rn(a5,s);a6=__fmul_rn(a6,s);a7=__fmul_rn(a7,s);}
    }
  } else {
    unsigned long long p0=pk(a0,a1),p1=pk(a2,a3),p2=pk(a4,a5),p3=pk(a6,a7), ss=pk(s,s);
    for(int i=0;i<iters;++i){
      #pragma unroll
      for(int u=0;u<8;++u){ p0=mul2(p0,ss);p1=mul2(p1,ss);p2=mul2(p2,ss);p3=mul2(p3,ss);}
    }
    upk(p0,a0,a1);upk(p1,a2,a3);upk(p2,a4,a5);upk(p3,a6,a7);
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=a0+a1+a2+a3+a4+a5+a6+a7;
}
int main(){
  float* d; cudaMalloc(&d, 148*8*256*4);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters=20000;
  for(int mode=0;mode<4;++mode){
    for(int rep=0;rep<2;++rep){
    cudaEventRecord(e0);
    if(mode==0)k<0><<<148*8,256>>>(d,iters,1.0001f); else if(mode==1)k<1><<<148*8,256>>>(d,iters,1.0001f); else if(mode==2)k<2><<<148*8,256>>>(d,iters,1.0001f); else k<3><<<148*8,256>>>(d,iters,1.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    double flops = 148.0*8*256*iters*64.0;
    if(rep) printf("mode %d: %.3f ms  %.2f T lane-ops/s\n", mode, ms, flops/ms*1e-9);
    }
  }
  return 0;
}
