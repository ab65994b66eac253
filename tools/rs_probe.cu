// micro-benchmark: the K2 resampler block loop (branch-free shift-register form) in isolation
#include <cstdio>
#include <cuda_runtime.h>
constexpr int kTcCh = 128, kTcN = 176, kTcBlocks = 11;
struct P { unsigned rs_emit[6]; int C; float g_out; };
__global__ void k(float *out, long long *cyc, P p, int reps, int nwarps_active)
{
    extern __shared__ float smem[];
    float *stage = smem;                 // [176][128]
    float *rs_seq = smem + 176 * 128;    // [176][16]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 176 * 128; i += blockDim.x) stage[i] = 0.001f * (i % 97);
    for (int i = threadIdx.x; i < 176 * 16; i += blockDim.x) rs_seq[i] = 0.01f * (i % 13);
    __syncthreads();
    if (warp >= nwarps_active) return;
    const float *st = stage + (warp & 3) * 32 + lane;
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; j++) acc[j] = 0.f;
    float *outp = out + (size_t)blockIdx.x * 147 * p.C * 0 + (warp & 3) * 32 + lane;
    const float g_out = p.g_out;
    long long c0 = clock64();
    for (int rep = 0; rep < reps; rep++) {
        outp = out + (warp & 3) * 32 + lane;
#pragma unroll 1
        for (int blk = 0; blk < kTcBlocks; blk++) {
            const int nrows = (blk == kTcBlocks - 1) ? 15 : 16;
            const unsigned emask = (p.rs_emit[blk >> 1] >> ((blk & 1) * 16)) & ((1u << nrows) - 1u);
            const float *col = st + 16 * blk * kTcCh;
            const float4 *cq = reinterpret_cast<const float4 *>(rs_seq + 16 * blk * 16);
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const float y = col[(i < nrows ? i : nrows - 1) * kTcCh];
                const float4 c0 = cq[4 * i], c1 = cq[4 * i + 1], c2 = cq[4 * i + 2], c3 = cq[4 * i + 3];
                const float cf[16] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x, c2.y, c2.z, c2.w, c3.x, c3.y, c3.z, c3.w};
                const bool emit = (emask >> i) & 1u;
                float tt[17];
#pragma unroll
                for (int j = 0; j < 16; j++) tt[j] = fmaf(cf[j], y, acc[j]);
                tt[16] = 0.f;
#pragma unroll
                for (int j = 0; j < 16; j++) acc[j] = emit ? tt[j + 1] : tt[j];
                if (emit) *outp = tt[0] * g_out;
                outp += emit ? p.C : 0;
            }
        }
    }
    long long c1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = c1 - c0;
    float s = 0;
    for (int j = 0; j < 16; j++) s += acc[j];
    out[(size_t)147 * p.C + blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    float *o; long long *c, h;
    P p; p.C = 1024; p.g_out = 1.f;
    for (int i = 0; i < 6; i++) p.rs_emit[i] = 0xFFFEFFFEu;
    cudaMalloc(&o, 4 * (147 * 1024 + 148 * 1024)); cudaMalloc(&c, 8);
    const int smem = (176 * 128 + 176 * 16) * 4;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int warps = 4; warps <= 16; warps *= 2) {
        k<<<148, 32 * warps, smem>>>(o, c, p, 20, warps);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("warps/SM %2d: %.0f cycles per tile (%.1f per row)\n", warps, (double)h / 20, (double)h / 20 / 175);
    }
    return 0;
}
