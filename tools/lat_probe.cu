// micro-benchmark: dependent-chain latency and throughput of DFMA / F2F / FFMA on this GPU
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, int n, double a, double b, float fa)
{
    double s = threadIdx.x * 1e-9, t0 = 1.0, t1 = 2.0, t2 = 3.0, t3 = 4.0;
    long long c0 = clock64();
    for (int i = 0; i < n; i++) s = fma(s, a, b);                       // dependent DFMA chain
    long long c1 = clock64();
    for (int i = 0; i < n; i++) { t0 = fma(t0, a, b); t1 = fma(t1, a, b); t2 = fma(t2, a, b); t3 = fma(t3, a, b); }  // 4 independent chains
    long long c2 = clock64();
    float f = (float)s;
    for (int i = 0; i < n; i++) { double d = (double)f; f = (float)(d * a); }  // F2F.F64.F32 -> DMUL -> F2F.F32.F64
    long long c3 = clock64();
    float g = f;
    for (int i = 0; i < n; i++) g = fmaf(g, fa, 1.0f);                   // dependent FFMA chain
    long long c4 = clock64();
    double u = s;
    for (int i = 0; i < n; i++) u = u + b;                                // dependent DADD
    long long c5 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = c1 - c0; cyc[1] = c2 - c1; cyc[2] = c3 - c2; cyc[3] = c4 - c3; cyc[4] = c5 - c4; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + t0 + t1 + t2 + t3 + f + g + u;
}
int main()
{
    double *o; long long *c, h[5];
    cudaMalloc(&o, 8 * 148 * 1024); cudaMalloc(&c, 40);
    const int n = 4096;
    for (int warps = 1; warps <= 16; warps *= 2) {
        k<<<148, 32 * warps>>>(o, c, n, 0.999, 1e-3, 0.999f);
        cudaDeviceSynchronize();
        cudaMemcpy(h, c, 40, cudaMemcpyDeviceToHost);
        printf("warps/SM %2d: DFMA dep %.1f cyc/op | 4 indep chains %.1f cyc/iter | f2f+dmul+f2f %.1f cyc/iter | FFMA dep %.1f | DADD dep %.1f\n",
               warps, (double)h[0] / n, (double)h[1] / n, (double)h[2] / n, (double)h[3] / n, (double)h[4] / n);
    }
    return 0;
}
