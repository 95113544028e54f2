// Dependent-issue latency of FP64 instructions on B200 (one warp, dependent chains), and issue rate with k independent chains.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP> __global__ void dfma_chain(double *out, int n, long long *cyc)
{
    double a[ILP];
    for (int i = 0; i < ILP; i++) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
    const double b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int k = 0; k < n; k++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) a[i] = fma(a[i], b, c);
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; i++) s += a[i];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void rcp_chain(double *out, int n, long long *cyc)
{
    double a = 1.5 + threadIdx.x * 1e-9;
    long long t0 = clock64();
    for (int k = 0; k < n; k++) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a)); a = y; }
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void lds_chain(double *out, int n, long long *cyc)
{
    __shared__ int idx[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i + 32) & 1023;
    __syncthreads();
    int j = threadIdx.x;
    long long t0 = clock64();
    for (int k = 0; k < n; k++) j = idx[j];
    long long t1 = clock64();
    out[threadIdx.x] = j;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void bar_chain(double *out, int n, long long *cyc)
{
    long long t0 = clock64();
    for (int k = 0; k < n; k++) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double *d; long long *c, h;
    cudaMalloc(&d, 1 << 20); cudaMalloc(&c, 8);
    const int n = 4096;
#define RUN(name, kern, grid, block, per) kern<<<grid, block>>>(d, n, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%-40s %.2f cycles per %s\n", name, (double)h / n, per);
    RUN("DFMA dependent, 1 warp", dfma_chain<1>, 1, 32, "op");
    RUN("DFMA ILP 2, 1 warp (per 2 ops)", dfma_chain<2>, 1, 32, "iter");
    RUN("DFMA ILP 4, 1 warp (per 4 ops)", dfma_chain<4>, 1, 32, "iter");
    RUN("DFMA ILP 8, 1 warp (per 8 ops)", dfma_chain<8>, 1, 32, "iter");
    RUN("DFMA ILP 1, 4 warps (1/SMSP)", dfma_chain<1>, 1, 128, "op");
    RUN("DFMA ILP 1, 16 warps (4/SMSP)", dfma_chain<1>, 1, 512, "op");
    RUN("DFMA ILP 4, 16 warps (4/SMSP, per 4)", dfma_chain<4>, 1, 512, "iter");
    RUN("MUFU.RCP64H dependent", rcp_chain, 1, 32, "op");
    RUN("LDS dependent", lds_chain, 1, 32, "op");
    RUN("__syncthreads, 9 warps", bar_chain, 1, 288, "barrier");
    RUN("__syncthreads, 2 warps", bar_chain, 1, 64, "barrier");
    return 0;
}
