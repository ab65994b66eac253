// micro-benchmark: FP64 pipe throughput / latency and f32<->f64 conversion cost on B200 (development probe for K2)
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double *out, long long *cyc, double a, double b, int reps)
{
    double acc[8];
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { acc[i] = threadIdx.x * 1e-3 + i; f[i] = (float)acc[i]; }
    long long c0 = clock64();
    for (int r = 0; r < reps; r++) {
        if (MODE == 0) {  // 8 independent DFMA chains
#pragma unroll
            for (int u = 0; u < 8; u++)
#pragma unroll
                for (int i = 0; i < 8; i++) acc[i] = fma(acc[i], a, b);
        } else if (MODE == 1) {  // one dependent chain
#pragma unroll
            for (int u = 0; u < 64; u++) acc[0] = fma(acc[0], a, b);
        } else if (MODE == 2) {  // f32 -> f64 -> f32 conversions, independent
#pragma unroll
            for (int u = 0; u < 8; u++)
#pragma unroll
                for (int i = 0; i < 8; i++) { double d = (double)f[i]; d += 1.0; f[i] = (float)d; }
        } else if (MODE == 3) {  // 4 chains
#pragma unroll
            for (int u = 0; u < 16; u++)
#pragma unroll
                for (int i = 0; i < 4; i++) acc[i] = fma(acc[i], a, b);
        } else if (MODE == 4) {  // 2 chains
#pragma unroll
            for (int u = 0; u < 32; u++)
#pragma unroll
                for (int i = 0; i < 2; i++) acc[i] = fma(acc[i], a, b);
        }
    }
    long long c1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = c1 - c0;
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    double *o; long long *c, h;
    cudaMalloc(&o, 8 * 148 * 1024); cudaMalloc(&c, 8);
    const int reps = 200;
    const char *names[] = {"8 indep DFMA chains", "1 dependent DFMA chain", "cvt f32->f64, DADD, cvt f64->f32 (8 indep)", "4 chains", "2 chains"};
    for (int mode = 0; mode < 5; mode++)
        for (int warps = 4; warps <= 16; warps *= 2) {
            if (mode == 0) k<0><<<148, 32 * warps>>>(o, c, 0.999, 0.5, reps);
            if (mode == 1) k<1><<<148, 32 * warps>>>(o, c, 0.999, 0.5, reps);
            if (mode == 2) k<2><<<148, 32 * warps>>>(o, c, 0.999, 0.5, reps);
            if (mode == 3) k<3><<<148, 32 * warps>>>(o, c, 0.999, 0.5, reps);
            if (mode == 4) k<4><<<148, 32 * warps>>>(o, c, 0.999, 0.5, reps);
            cudaDeviceSynchronize();
            cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            printf("%-45s warps/SM %2d: %.2f cycles per warp-level op group (64 ops/rep) -> %.2f cyc/op/warp\n", names[mode], warps,
                   (double)h / reps, (double)h / reps / 64);
        }
    return 0;
}
