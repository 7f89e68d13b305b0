// Does the driver give device memory back when a green context (and its streams) is destroyed?  Free memory after cycles of
// (a) cuGreenCtxCreate/Destroy, (b) + cuGreenCtxStreamCreate/cuStreamDestroy, (c) + one kernel launched into the stream.
// nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/greenctx_leak_probe tools/greenctx_leak_probe.cu -lcuda && /tmp/greenctx_leak_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
__global__ void k(int *p) { if (p) *p = 1; }
#define CK(x) do { CUresult r = (x); if (r != CUDA_SUCCESS) { printf("%s -> %d\n", #x, (int)r); return 1; } } while (0)
int main() {
    cudaFree(0);
    CUdevice dev; CK(cuDeviceGet(&dev, 0));
    int *d; cudaMalloc(&d, 4);
    for (int mode = 0; mode < 3; ++mode) {
        size_t f0 = 0, f1 = 0, tot;
        for (int it = 0; it < 10; ++it) {
            if (it == 2) { cudaDeviceSynchronize(); cudaMemGetInfo(&f0, &tot); }
            CUdevResource all, back, front;
            CK(cuDeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
            unsigned int groups = 1;
            CK(cuDevSmResourceSplitByCount(&back, &groups, &all, &front, 0, 32));
            CUdevResourceDesc db, df;
            CK(cuDevResourceGenerateDesc(&db, &back, 1));
            CK(cuDevResourceGenerateDesc(&df, &front, 1));
            CUgreenCtx gb, gf;
            CK(cuGreenCtxCreate(&gb, db, dev, CU_GREEN_CTX_DEFAULT_STREAM));
            CK(cuGreenCtxCreate(&gf, df, dev, CU_GREEN_CTX_DEFAULT_STREAM));
            CUstream sb = nullptr, sf = nullptr;
            if (mode >= 1) {
                CK(cuGreenCtxStreamCreate(&sb, gb, CU_STREAM_NON_BLOCKING, 0));
                CK(cuGreenCtxStreamCreate(&sf, gf, CU_STREAM_NON_BLOCKING, 0));
            }
            if (mode >= 2) { k<<<1, 32, 0, (cudaStream_t)sb>>>(d); k<<<1, 32, 0, (cudaStream_t)sf>>>(d); cudaDeviceSynchronize(); }
            if (sb) CK(cuStreamDestroy(sb));
            if (sf) CK(cuStreamDestroy(sf));
            CK(cuGreenCtxDestroy(gb));
            CK(cuGreenCtxDestroy(gf));
        }
        cudaDeviceSynchronize(); cudaMemGetInfo(&f1, &tot);
        printf("mode %d: %.2f MiB per cycle (8 cycles)\n", mode, (double)((long long)f0 - (long long)f1) / 8 / (1 << 20));
    }
    return 0;
}
