// Does an FP64 instruction block the issue port for 2 cycles?  DFMA interleaved with independent integer ops.
#include <cstdio>
#include <cuda_runtime.h>
template <int NI> __global__ void mix(double *out, int n, long long *cyc)
{
    double a[4];
    for (int i = 0; i < 4; i++) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
    unsigned u[8];
    for (int i = 0; i < 8; i++) u[i] = threadIdx.x * 7 + i;
    const double b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int k = 0; k < n; k++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            a[i] = fma(a[i], b, c);
#pragma unroll
            for (int j = 0; j < NI; j++) u[(i * NI + j) & 7] = u[(i * NI + j) & 7] * 3u + 12345u + k;   // IMAD
        }
    }
    long long t1 = clock64();
    double s = 0; unsigned v = 0;
    for (int i = 0; i < 4; i++) s += a[i];
    for (int i = 0; i < 8; i++) v ^= u[i];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s + v;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double *d; long long *c, h;
    cudaMalloc(&d, 1 << 20); cudaMalloc(&c, 8);
    const int n = 4096;
#define RUN(name, kern, grid, block) kern<<<grid, block>>>(d, n, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%-50s %.2f cycles per DFMA (per SMSP)\n", name, (double)h / n / 4 / ((block) / 128.0));
    RUN("16 warps: DFMA only", mix<0>, 1, 512);
    RUN("16 warps: DFMA + 1 int", mix<1>, 1, 512);
    RUN("16 warps: DFMA + 2 int", mix<2>, 1, 512);
    RUN("16 warps: DFMA + 3 int", mix<3>, 1, 512);
    return 0;
}
