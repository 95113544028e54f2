// Accuracy of the FP64 reciprocal seed (MUFU.RCP64H via rcp.approx.ftz.f64) and of refinement sequences.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ double seed(double b) { double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b)); return y; }
__global__ void k(double lo, double hi, int n, double *out)
{
    double m0 = 0, m1 = 0, m2 = 0, m3 = 0, m4 = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double b = lo + (hi - lo) * ((double)i + 0.37) / n;
        const double ex = 1.0 / b;
        double y = seed(b);
        m0 = fmax(m0, fabs(y - ex) / ex);
        // one cubic step
        double e = fma(-b, y, 1.0);
        double t = fma(e, e, e);
        double yc = fma(y, t, y);
        m1 = fmax(m1, fabs(yc - ex) / ex);
        // cubic + quadratic
        double e2 = fma(-b, yc, 1.0);
        double ycq = fma(yc, e2, yc);
        m2 = fmax(m2, fabs(ycq - ex) / ex);
        // two quadratic
        double y1 = fma(y, e, y);
        double e1 = fma(-b, y1, 1.0);
        double y2 = fma(y1, e1, y1);
        m3 = fmax(m3, fabs(y2 - ex) / ex);
        // three quadratic (current)
        double e3 = fma(-b, y2, 1.0);
        double y3 = fma(y2, e3, y2);
        m4 = fmax(m4, fabs(y3 - ex) / ex);
    }
    double v[5] = {m0, m1, m2, m3, m4};
    for (int j = 0; j < 5; j++) {
        unsigned long long *p = (unsigned long long *)(out + j);
        atomicMax(p, __double_as_longlong(v[j]));
    }
}
int main()
{
    double *d, h[5];
    cudaMalloc(&d, sizeof h);
    const double ranges[3][2] = {{0.5, 2.0}, {0.9, 1.1}, {1e-3, 1e3}};
    for (auto &r : ranges) {
        cudaMemset(d, 0, sizeof h);
        k<<<592, 256>>>(r[0], r[1], 1 << 28, d);
        cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
        printf("[%g, %g]: seed %.3e  cubic %.3e  cubic+quad %.3e  2 quad %.3e  3 quad %.3e  (eps = %.3e)\n", r[0], r[1], h[0], h[1], h[2], h[3], h[4], ldexp(1.0, -53));
    }
    return 0;
}
